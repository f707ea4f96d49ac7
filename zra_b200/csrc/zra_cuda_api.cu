// zra_cuda_api.cu — the additive device-pointer C ABI declared in include/zra_b200.h.
#include <cstring>

#include "../../include/zra_b200.h"
#include "gpu_context.h"
#include "zra_format.h"

using namespace zrab;

struct ZraCudaContext {
  GpuContext gpu;
  explicit ZraCudaContext(int device) : gpu(device) {}
};

namespace {
  ZraStatus st(ZraStatusCode z, int zstd = 0) { return ZraStatus{z, zstd}; }
  ZraStatus from(const DecodeResult& r) {
    if (r.cudaFailed) return st(ZStdError, 1);
    if (r.zstd) return st(ZStdError, r.zstd);
    return st(Success);
  }
  static_assert(sizeof(ZraCudaFrame) == sizeof(HostFrame), "frame descriptor layouts must agree");

  // Reads and validates the fixed header of a device-resident archive.
  ZraStatus read_info(GpuContext& g, const void* dArchive, size_t n, ArchiveInfo* info, cudaStream_t s) {
    if (n <= kFixedHeaderSize) return st(OutOfBoundsAccess);
    uint8_t raw[kFixedHeaderSize];
    if (g.check(cudaMemcpyAsync(raw, dArchive, sizeof(raw), cudaMemcpyDeviceToHost, s), "header readback") ||
        g.check(cudaStreamSynchronize(s), "header readback"))
      return st(ZStdError, 1);
    FixedHeaderFields f = parse_fixed_header(raw);
    if (f.magic != kZraMagic || f.version > kZraVersion) return st(HeaderInvalid);
    if (f.version != 1) return st(ZraVersionLow);
    info->headerSize = f.headerSize + 8;
    info->tableSize = f.tableSize;
    info->frameSize = f.frameSize;
    info->metaSize = f.metaSize;
    info->uncompressedSize = f.uncompressedSize;
    info->frames = f.tableSize ? f.tableSize - 1 : 0;
    if (n < info->headerSize) return st(OutOfBoundsAccess);
    if (!f.tableSize || (info->frames && !f.frameSize)) return st(HeaderInvalid);
    if (f.frameSize && f.tableSize != table_entries(f.uncompressedSize, f.frameSize)) return st(HeaderInvalid);
    if ((uint64_t)kFixedHeaderSize + f.metaSize + kEntrySize * (uint64_t)f.tableSize != info->headerSize) return st(HeaderInvalid);
    return st(Success);
  }
}  // namespace

extern "C" {

ZraStatus ZraCudaCreateContext(ZraCudaContext** context, int device) {
  auto* c = new ZraCudaContext(device);
  *context = c;  // returned even on failure so that ZraCudaGetLastError can explain it
  return c->gpu.ok() ? st(Success) : st(ZStdError, 1);
}

void ZraCudaDestroyContext(ZraCudaContext* context) { delete context; }

const char* ZraCudaGetLastError(ZraCudaContext* context) { return context->gpu.last_error().c_str(); }

uint64_t ZraCudaGetLaunchCount(ZraCudaContext* context) { return context->gpu.launches(); }

void ZraCudaSetProfiling(ZraCudaContext* context, int enabled) {
  context->gpu.set_profiling(enabled != 0);
  context->gpu.timer.reset();
}

int ZraCudaGetKernelProfile(ZraCudaContext* context, int index, const char** name, double* totalMs, uint64_t* launches) {
  if (index < 1 || index >= K_COUNT) return 0;
  *name = kernel_name(index);
  *totalMs = context->gpu.timer.ms[index];
  *launches = context->gpu.timer.launches[index];
  return 1;
}

ZraStatus ZraCudaDecodeFrames(ZraCudaContext* context, const void* dSrc, size_t srcSize, const ZraCudaFrame* frames, uint32_t count,
                              void* dDst, uint32_t* frameSizes, uint32_t* failedFrame, void* stream) {
  GpuContext& g = context->gpu;
  if (!g.ok()) return st(ZStdError, 1);
  if (reinterpret_cast<uintptr_t>(dSrc) & 15u) { g.fail("zra-b200: device source pointer must be 16-byte aligned"); return st(ZStdError, 1); }
  uint32_t maxCap = 0;
  for (uint32_t i = 0; i < count; i++) maxCap = frames[i].dstCapacity > maxCap ? frames[i].dstCapacity : maxCap;
  DecodeResult r = g.decode(dSrc, srcSize, reinterpret_cast<const HostFrame*>(frames), nullptr, 0, count, maxCap, dDst, frameSizes,
                            static_cast<cudaStream_t>(stream));
  if (failedFrame) *failedFrame = r.failedFrame;
  return from(r);
}

ZraStatus ZraCudaDecompressFrames(ZraCudaContext* context, const void* dArchive, size_t archiveSize, uint64_t firstFrame,
                                  uint64_t frameCount, void* dOutput, size_t outputCapacity, void* stream) {
  GpuContext& g = context->gpu;
  if (!g.ok()) return st(ZStdError, 1);
  g.bind();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (reinterpret_cast<uintptr_t>(dArchive) & 15u) { g.fail("zra-b200: device archive pointer must be 16-byte aligned"); return st(ZStdError, 1); }
  ArchiveInfo info;
  ZraStatus hs = read_info(g, dArchive, archiveSize, &info, s);
  if (hs.zra != Success) return hs;
  if (firstFrame > info.frames || frameCount > info.frames - firstFrame) return st(OutOfBoundsAccess);
  if (!frameCount) return st(Success);
  uint64_t begin = firstFrame * info.frameSize;
  uint64_t end = firstFrame + frameCount == info.frames ? info.uncompressedSize : (firstFrame + frameCount) * (uint64_t)info.frameSize;
  if (outputCapacity < end - begin) return st(OutputBufferTooSmall);
  uint32_t maxCap = (uint32_t)(info.frameSize < info.uncompressedSize ? info.frameSize : info.uncompressedSize);
  return from(g.decode(dArchive, archiveSize, nullptr, &info, firstFrame, frameCount, maxCap, dOutput, nullptr, s));
}

ZraStatus ZraCudaDecompressBuffer(ZraCudaContext* context, const void* dArchive, size_t archiveSize, void* dOutput,
                                  size_t outputCapacity, void* stream) {
  GpuContext& g = context->gpu;
  if (!g.ok()) return st(ZStdError, 1);
  g.bind();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (reinterpret_cast<uintptr_t>(dArchive) & 15u) { g.fail("zra-b200: device archive pointer must be 16-byte aligned"); return st(ZStdError, 1); }
  ArchiveInfo info;
  ZraStatus hs = read_info(g, dArchive, archiveSize, &info, s);
  if (hs.zra != Success) return hs;
  if (outputCapacity < info.uncompressedSize) return st(OutputBufferTooSmall);
  if (!info.frames) return st(Success);
  uint32_t maxCap = (uint32_t)(info.frameSize < info.uncompressedSize ? info.frameSize : info.uncompressedSize);
  return from(g.decode(dArchive, archiveSize, nullptr, &info, 0, info.frames, maxCap, dOutput, nullptr, s));
}

ZraStatus ZraCudaDecompressRABatch(ZraCudaContext* context, const void* dArchive, size_t archiveSize, const uint64_t* dOffsets,
                                   const uint32_t* dSizes, const uint64_t* dOutOffsets, uint32_t uniformSize, uint32_t maxSize,
                                   uint64_t count, void* dOutput, uint64_t* uniqueFrames, uint64_t* badRequest, void* stream) {
  GpuContext& g = context->gpu;
  if (!g.ok()) return st(ZStdError, 1);
  g.bind();
  if (uniqueFrames) *uniqueFrames = 0;
  if (badRequest) *badRequest = ~0ull;
  if (!dSizes) maxSize = uniformSize;
  if (dSizes && !dOutOffsets) { g.fail("zra-b200: per-request sizes need per-request output offsets"); return st(ZStdError, 42); }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (reinterpret_cast<uintptr_t>(dArchive) & 15u) { g.fail("zra-b200: device archive pointer must be 16-byte aligned"); return st(ZStdError, 1); }
  ArchiveInfo info;
  ZraStatus hs = read_info(g, dArchive, archiveSize, &info, s);
  if (hs.zra != Success) return hs;
  RaBatch b{};
  b.offsets = dOffsets; b.sizes = dSizes; b.outOffsets = dOutOffsets; b.uniformSize = uniformSize;
  GpuContext::RaResult r = g.random_access(dArchive, archiveSize, info, b, count, maxSize, dOutput, s);
  if (uniqueFrames) *uniqueFrames = r.uniqueFrames;
  if (badRequest) *badRequest = r.badRequest;
  if (r.cudaFailed) return st(ZStdError, 1);
  if (r.zra) return st(static_cast<ZraStatusCode>(r.zra), r.zstd);
  return st(Success);
}

ZraStatus ZraCudaCompressBuffer(ZraCudaContext* context, const void* dInput, size_t inputSize, void* dOutput, size_t outputCapacity,
                                size_t* outputSize, int8_t compressionLevel, uint32_t frameSize, bool checksum, const void* metaBuffer,
                                size_t metaSize, void* stream) {
  GpuContext& g = context->gpu;
  if (!g.ok()) return st(ZStdError, 1);
  if (!frameSize) return st(ZStdError, 42);  // parameter_outOfBound
  uint32_t table = table_entries(inputSize, frameSize);
  size_t need = kFixedHeaderSize + metaSize + kEntrySize * (size_t)table + zstd_compress_bound(frameSize) * (size_t)(table - 1);
  if (outputCapacity < need) return st(OutputBufferTooSmall);
  GpuContext::CompressStatus r = g.compress_archive(dInput, inputSize, dOutput, outputCapacity, compressionLevel, frameSize, checksum,
                                                    static_cast<const uint8_t*>(metaBuffer), metaSize, /*refMetaQuirk=*/false,
                                                    static_cast<cudaStream_t>(stream));
  if (r.cudaFailed) return st(ZStdError, 1);
  if (r.zra) return st(static_cast<ZraStatusCode>(r.zra));
  *outputSize = r.total;
  return st(Success);
}

ZraStatus ZraCudaCompressFrames(ZraCudaContext* context, const void* dInput, size_t inputSize, uint32_t frameSize,
                                int8_t compressionLevel, bool checksum, void* dOutput, size_t outputCapacity, uint64_t* frameSizes,
                                size_t* outputSize, void* stream) {
  GpuContext& g = context->gpu;
  if (!g.ok()) return st(ZStdError, 1);
  if (!frameSize) return st(ZStdError, 42);  // parameter_outOfBound
  *outputSize = 0;
  GpuContext::CompressStatus r = g.compress_frames(dInput, inputSize, frameSize, compressionLevel, checksum, dOutput, outputCapacity,
                                                   frameSizes, static_cast<cudaStream_t>(stream));
  if (r.cudaFailed) return st(ZStdError, 1);
  if (r.zra) return st(static_cast<ZraStatusCode>(r.zra));
  *outputSize = r.total;
  return st(Success);
}

size_t ZraShardHeaderSize(uint64_t frames, size_t metaSize) { return kFixedHeaderSize + metaSize + kEntrySize * (size_t)(frames + 1); }

ZraStatus ZraShardBuildHeader(uint64_t uncompressedSize, uint32_t frameSize, const void* metaBuffer, size_t metaSize,
                              const uint64_t* frameSizes, uint64_t frames, void* out, size_t outCapacity) {
  const size_t total = ZraShardHeaderSize(frames, metaSize);
  if (outCapacity < total) return st(OutputBufferTooSmall);
  if (!frameSize || frames + 1 != table_entries(uncompressedSize, frameSize)) return st(InputFrameSizeMismatch);
  uint8_t* p = static_cast<uint8_t*>(out);
  write_fixed_header(p, uncompressedSize, (uint32_t)(frames + 1), frameSize, (uint32_t)metaSize);
  if (metaSize) memcpy(p + kFixedHeaderSize, metaBuffer, metaSize);
  uint8_t* table = p + kFixedHeaderSize + metaSize;
  uint64_t offset = 0;
  for (uint64_t f = 0; f < frames; f++) {
    put_le(table + kEntrySize * f, offset, 5);
    offset += frameSizes[f];
    if (offset >= kMaxCompressedSize) return st(CompressedSizeTooLarge);
  }
  put_le(table + kEntrySize * frames, offset, 5);
  put_le(p + 14, header_hash_host(p, total), 4);
  return st(Success);
}

ZraStatus ZraVerifyHeaderCrc(const void* archive, size_t size) {
  const uint8_t* p = static_cast<const uint8_t*>(archive);
  if (!p || size < kFixedHeaderSize) return st(OutOfBoundsAccess);
  if (get_le(p + 8, 4) != kZraMagic || get_le(p + 12, 2) > kZraVersion) return st(HeaderInvalid);
  const uint64_t total = get_le(p + 4, 4) + 8;
  const uint64_t expect = (uint64_t)kFixedHeaderSize + get_le(p + 34, 4) + kEntrySize * get_le(p + 26, 4);
  if (total != expect) return st(HeaderInvalid);  // the three size fields must agree with each other
  if (total > size) return st(OutOfBoundsAccess);
  return st(header_hash_host(p, (size_t)total) == (uint32_t)get_le(p + 14, 4) ? Success : HeaderInvalid);
}

}  // extern "C"
