// host_ops.cu — host-buffer operations: H2D staging, kernel launches, D2H of results.
#include "host_ops.h"

#include <algorithm>
#include <cstring>
#include <vector>

#include "zra_format.h"

namespace zrab {

namespace {
  OpStatus cuda_failed() {
    OpStatus s;
    s.cuda = true;
    return s;
  }
  OpStatus from_decode(const DecodeResult& r) {
    OpStatus s;
    if (r.cudaFailed) s.cuda = true;
    else if (r.zstd) { s.zra = 1; s.zstd = r.zstd; }
    return s;
  }
  OpStatus zra_error(int code, int zstd = 0) {
    OpStatus s;
    s.zra = code;
    s.zstd = zstd;
    return s;
  }
  size_t pad4(size_t n) { return (n + 3) & ~size_t(3); }
}  // namespace

OpStatus host_decompress_archive(GpuContext* g, const uint8_t* archive, size_t n, const ArchiveInfo& info, uint8_t* out) {
  // geometry the frame-parallel decoder relies on (the reference's serial decoder never reads the table)
  if (!info.tableSize) return zra_error(3);
  if (info.frames && !info.frameSize) return zra_error(3);
  if (info.frameSize && info.tableSize != table_entries(info.uncompressedSize, info.frameSize)) return zra_error(3);
  if (n < info.headerSize) return zra_error(5);
  // the seek table is read below: it must lie where the header says, inside the header (a crafted metaSize would
  // otherwise send the reads far outside the caller's buffer); same rule as the device path (zra_cuda_api.cu read_info)
  if (38ull + info.metaSize + kEntrySize * (uint64_t)info.tableSize != info.headerSize) return zra_error(3);
  uint64_t lastEntry = get_le(archive + 38 + info.metaSize + kEntrySize * (size_t)(info.tableSize - 1), 5);
  if ((uint64_t)info.headerSize + lastEntry != n) return zra_error(1, 72);  // truncated or trailing bytes: srcSize_wrong
  if (!info.frames) return OpStatus{};
  cudaStream_t st = g->stream();
  uint8_t* dIn = static_cast<uint8_t*>(g->ensure(g->stageIn, pad4(n) + 16));
  uint8_t* dOut = static_cast<uint8_t*>(g->ensure(g->stageOut, info.uncompressedSize + 16));
  if (!dIn || !dOut) return cuda_failed();
  // descriptors from the host copy of the seek table; the chunk pipeline uploads each chunk's
  // compressed range, decodes it and downloads its output while other chunks are in flight
  const uint8_t* table = archive + 38 + info.metaSize;
  std::vector<HostFrame> frames(info.frames);
  uint64_t prev = get_le(table, 5);
  for (uint64_t f = 0; f < info.frames; f++) {
    uint64_t next = get_le(table + kEntrySize * (f + 1), 5);
    HostFrame& d = frames[f];
    uint64_t begin = f * info.frameSize;
    d.srcOff = info.headerSize + prev;
    d.dstOff = begin;
    d.dstCap = (uint32_t)std::min<uint64_t>(info.frameSize, info.uncompressedSize - begin);
    d.exact = 1;
    d.pad = 0;
    // an inconsistent entry decodes as a zero-length frame -> srcSize_wrong, like the device table reader
    d.srcLen = (next < prev || info.headerSize + next > n || next - prev > 0xFFFFFFFFull) ? 0 : (uint32_t)(next - prev);
    if (!d.srcLen) d.srcOff = info.headerSize;
    prev = next;
  }
  uint32_t maxCap = (uint32_t)std::min<uint64_t>(info.frameSize, info.uncompressedSize);
  HostStaging io;
  io.hostSrc = archive;
  io.hostDst = out;
  DecodeResult r = g->decode(dIn, n, frames.data(), nullptr, 0, info.frames, maxCap, dOut, nullptr, st, &io);
  return from_decode(r);
}

OpStatus host_decode_frames(GpuContext* g, const uint8_t* src, size_t srcSize, const HostFrame* frames, size_t nFrames,
                            uint32_t frameSize, uint64_t skip, uint64_t size, uint8_t* out, const std::function<void()>* whileBusy) {
  if (!nFrames) return OpStatus{};
  cudaStream_t st = g->stream();
  uint64_t staged = 0;
  uint32_t maxCap = 0;
  for (size_t i = 0; i < nFrames; i++) {
    staged = std::max<uint64_t>(staged, frames[i].dstOff + frames[i].dstCap);
    maxCap = std::max(maxCap, frames[i].dstCap);
  }
  (void)frameSize;
  uint8_t* dIn = static_cast<uint8_t*>(g->ensure(g->stageIn, pad4(srcSize) + 16));
  uint8_t* dOut = static_cast<uint8_t*>(g->ensure(g->stageOut, staged + 16));
  if (!dIn || !dOut) return cuda_failed();
  if (skip + size > staged) return zra_error(5);
  HostStaging io;
  io.hostSrc = src;
  io.hostDst = out;
  io.dstSkip = skip;
  io.dstSize = size;
  io.whileBusy = whileBusy;
  return from_decode(g->decode(dIn, srcSize, frames, nullptr, 0, nFrames, maxCap, dOut, nullptr, st, &io));
}

OpStatus host_decompress_range(GpuContext* g, const uint8_t* archive, size_t n, const ArchiveInfo& info, uint64_t offset,
                               uint64_t size, uint8_t* out) {
  if (!info.frameSize) return zra_error(3);
  if (!size) return OpStatus{};
  if (n < info.headerSize) return zra_error(5);
  if (38ull + info.metaSize + kEntrySize * (uint64_t)info.tableSize != info.headerSize) return zra_error(3);  // (see above)
  const uint8_t* table = archive + 38 + info.metaSize;
  uint64_t first = offset / info.frameSize, last = (offset + size - 1) / info.frameSize + 1;
  if (last >= info.tableSize) return zra_error(5);
  uint64_t a = get_le(table + kEntrySize * first, 5), b = get_le(table + kEntrySize * last, 5);
  if (b < a || info.headerSize + b > n) return zra_error(1, 72);
  std::vector<HostFrame> frames(last - first);
  for (uint64_t f = first; f < last; f++) {
    HostFrame& d = frames[f - first];
    uint64_t fa = get_le(table + kEntrySize * f, 5), fb = get_le(table + kEntrySize * (f + 1), 5);
    if (fb < fa || fb > b || fb - fa > 0xFFFFFFFFull) return zra_error(1, 72);
    uint64_t begin = f * info.frameSize;
    d.srcOff = fa - a;
    d.srcLen = (uint32_t)(fb - fa);
    d.dstOff = begin - first * info.frameSize;
    d.dstCap = (uint32_t)std::min<uint64_t>(info.frameSize, info.uncompressedSize - begin);
    d.exact = 1;
    d.pad = 0;
  }
  return host_decode_frames(g, archive + info.headerSize + a, b - a, frames.data(), frames.size(), info.frameSize,
                            offset - first * info.frameSize, size, out);
}

OpStatus host_compress_buffer(GpuContext* g, const uint8_t* in, size_t n, uint8_t* out, size_t outCap, size_t* written, int level,
                              uint32_t frameSize, bool checksum, const uint8_t* meta, size_t metaSize) {
  cudaStream_t st = g->stream();
  uint8_t* dIn = static_cast<uint8_t*>(g->ensure(g->stageIn, pad4(n) + 64));
  uint8_t* dOut = static_cast<uint8_t*>(g->ensure(g->stageOut, outCap + 64));
  if (!dIn || !dOut) return cuda_failed();
  // zero the slack the 32-bit readers may touch past the end of the input
  if (g->check(cudaMemsetAsync(dIn + n, 0, pad4(n) + 64 - n, st), "memset")) return cuda_failed();
  // the input goes up, the frames are compressed and the archive comes down in overlapping batches (gpu_compress.cu)
  GpuContext::CompressStatus r = g->compress_archive(dIn, n, dOut, outCap, level, frameSize, checksum, meta, metaSize,
                                                     /*refMetaQuirk=*/true, st, in, out);
  if (r.cudaFailed) return cuda_failed();
  if (r.zra) return zra_error(r.zra);
  *written = r.total;
  return OpStatus{};
}

OpStatus host_compress_frames(GpuContext* g, const uint8_t* in, size_t n, uint32_t frameSize, int level, bool checksum, uint8_t* out,
                              size_t outCap, uint64_t* sizes, size_t* produced) {
  *produced = 0;
  if (!n) return OpStatus{};
  // The frames of one call are a complete archive without metadata: the archive path (upload, kernels and download
  // overlapped in batches) builds it in device memory, the frames go straight to `out`, and the seek table that stays
  // on the device gives the frame sizes.
  cudaStream_t st = g->stream();
  const uint64_t frames = (n + frameSize - 1) / frameSize;
  const size_t tableOff = kFixedHeaderSize, framesOff = tableOff + kEntrySize * (size_t)(frames + 1);
  uint8_t* dIn = static_cast<uint8_t*>(g->ensure(g->stageIn, pad4(n) + 64));
  uint8_t* dOut = static_cast<uint8_t*>(g->ensure(g->stageOut, framesOff + outCap + 64));
  if (!dIn || !dOut) return cuda_failed();
  if (g->check(cudaMemsetAsync(dIn + n, 0, pad4(n) + 64 - n, st), "memset")) return cuda_failed();
  GpuContext::CompressStatus r = g->compress_archive(dIn, n, dOut, framesOff + outCap, level, frameSize, checksum, nullptr, 0, false, st, in,
                                                     nullptr, out);
  if (r.cudaFailed) return cuda_failed();
  if (r.zra) return zra_error(r.zra);
  std::vector<uint8_t> table(kEntrySize * (size_t)(frames + 1));
  if (g->check(cudaMemcpyAsync(table.data(), dOut + tableOff, table.size(), cudaMemcpyDeviceToHost, st), "table download") ||
      g->check(cudaStreamSynchronize(st), "table download"))
    return cuda_failed();
  uint64_t prev = get_le(table.data(), 5);
  for (uint64_t f = 0; f < frames; f++) {
    const uint64_t next = get_le(table.data() + kEntrySize * (size_t)(f + 1), 5);
    sizes[f] = next - prev;
    prev = next;
  }
  *produced = r.total - framesOff;
  return OpStatus{};
}

}  // namespace zrab
