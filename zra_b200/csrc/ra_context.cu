// ra_context.cu — host driver of batched random access (GpuContext::random_access).
#include <algorithm>

#include "gpu_context.h"

namespace zrab {

GpuContext::RaResult GpuContext::random_access(const void* dArchive, size_t archiveSize, const ArchiveInfo& info, RaBatch b,
                                               uint64_t count, uint32_t maxSize, void* dOut, cudaStream_t st) {
  RaResult res;
  if (!count) return res;
  bind();
  auto fail_cuda = [&]() { res.cudaFailed = true; return res; };
  if (!info.frameSize || !info.frames) {  // nothing can be in bounds except empty reads
    res.zra = 5;
    res.badRequest = 0;
    return res;
  }
  b.frameSize = info.frameSize;
  b.uncompressedSize = info.uncompressedSize;
  if (!raHost_ && check(cudaMallocHost(&raHost_, 16), "cudaMallocHost")) return fail_cuda();
  // frame -> slot map, "empty" everywhere between batches
  if (raSlotFrames_ < info.frames || !raSlotOf.p) {
    raSlotFrames_ = 0;
    if (!ensure(raSlotOf, sizeof(uint32_t) * info.frames)) return fail_cuda();
    if (check(cudaMemsetAsync(raSlotOf.p, 0xFF, sizeof(uint32_t) * info.frames, st), "slot map init")) return fail_cuda();
    raSlotFrames_ = info.frames;
  }
  // sub-batches sized so that the decoded slots and the decode scratch fit the budget
  const uint64_t touch = (uint64_t)(maxSize ? (maxSize - 1) / info.frameSize : 0) + 2;  // frames one request can touch
  DecodeLayout one;
  const uint32_t cap = (uint32_t)std::min<uint64_t>(info.frameSize, info.uncompressedSize);
  const size_t perFrame = decode_scratch_bytes(1, cap, &one) + info.frameSize + 64;
  const uint64_t budget = 16ull << 30;
  const uint64_t budgetFrames = std::max<uint64_t>(touch, budget / perFrame);
  // a batch can never touch more than every frame of the archive: if those fit, the whole batch is one
  // sub-batch (frames are then decoded exactly once); otherwise bound the frames a sub-batch can touch
  const uint64_t perBatch = info.frames <= budgetFrames ? std::min<uint64_t>(count, 1u << 30)
                                                        : std::min<uint64_t>(std::max<uint64_t>(1, budgetFrames / touch), 1u << 30);
  uint32_t* slotOf = static_cast<uint32_t*>(raSlotOf.p);
  for (uint64_t r0 = 0; r0 < count; r0 += perBatch) {
    const uint32_t n = (uint32_t)std::min<uint64_t>(perBatch, count - r0);
    const uint64_t worst = std::min<uint64_t>((uint64_t)n * touch, info.frames);
    uint32_t* unique = static_cast<uint32_t*>(ensure(raUnique, sizeof(uint32_t) * worst + 16));
    if (!unique) return fail_cuda();
    uint32_t* counters = unique + worst;  // {unique frames, first bad request}: the 8 bytes after the list
    counters = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(counters) + 7) & ~uintptr_t(7));
    launch_ra_map(b, r0, n, slotOf, unique, counters, st);
    launches_ += 1;
    if (check(cudaMemcpyAsync(raHost_, counters, 8, cudaMemcpyDeviceToHost, st), "ra counters") ||
        check(cudaStreamSynchronize(st), "ra map"))
      return fail_cuda();
    const uint32_t nUnique = raHost_[0];
    auto reset = [&]() {
      launch_ra_reset(slotOf, unique, nUnique, st);
      launches_ += 1;
    };
    if (raHost_[1] != 0xFFFFFFFFu) {
      reset();
      cudaStreamSynchronize(st);
      res.zra = 5;  // OutOfBoundsAccess
      res.badRequest = r0 + raHost_[1];
      return res;
    }
    if (nUnique) {
      void* descs = ensure(raDescs, sizeof(HostFrame) * (size_t)nUnique);
      void* slots = ensure(raFrames, (size_t)nUnique * info.frameSize + 64);
      if (!descs || !slots) { reset(); return fail_cuda(); }
      launch_ra_descs(dArchive, 38ull + info.metaSize, info.headerSize, archiveSize, info.uncompressedSize, info.frameSize, unique,
                      nUnique, descs, st);
      launches_ += 1;
      DecodeResult d = decode(dArchive, archiveSize, static_cast<const HostFrame*>(descs), nullptr, 0, nUnique, cap, slots, nullptr, st);
      if (d.cudaFailed || d.zstd) {
        reset();
        cudaStreamSynchronize(st);
        res.cudaFailed = d.cudaFailed;
        if (d.zstd) { res.zra = 1; res.zstd = d.zstd; }
        return res;
      }
      launch_ra_gather(b, r0, n, slotOf, slots, dOut, st);
      launches_ += 1;
      reset();
      res.uniqueFrames += nUnique;
    }
  }
  if (check(cudaStreamSynchronize(st), "ra gather")) return fail_cuda();
  return res;
}

}  // namespace zrab
