// gpu_compress.cu — batched compression driver: frames -> encoder kernels -> scan -> seek table -> CRC.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "encode_launch.h"
#include "gpu_context.h"
#include "zra_format.h"

namespace zrab {

namespace {
size_t enc_budget() {
  static size_t v = [] {
    const char* s = getenv("ZRA_B200_ENC_SCRATCH_MB");
    size_t mb = s ? strtoull(s, nullptr, 10) : 12288;
    return std::max<size_t>(mb, 64) << 20;
  }();
  return v;
}
}  // namespace

GpuContext::CompressStatus GpuContext::compress_frames(const void* dIn, size_t n, uint32_t frameSize, int level, bool checksum,
                                                       void* dOut, size_t outCap, uint64_t* sizesHost, cudaStream_t st) {
  CompressStatus r;
  if (!n) return r;
  bind();
  const uint64_t frames = (n + frameSize - 1) / frameSize;
  const uint32_t lastLen = (uint32_t)(n - (frames - 1) * frameSize);
  EncodeLayout one;
  size_t perFrame = encode_scratch_bytes(1, frameSize, lastLen, level, &one);
  uint64_t batch = std::max<uint64_t>(1, enc_budget() / perFrame);
  batch = std::min<uint64_t>(batch, frames);
  uint64_t* dTotal = static_cast<uint64_t*>(ensure(misc, 4096 + 8 * (size_t)batch));
  if (!dTotal) { r.cudaFailed = true; return r; }
  uint64_t* dSizes64 = dTotal + 8;
  uint64_t produced = 0;
  for (uint64_t f0 = 0; f0 < frames; f0 += batch) {
    uint32_t nb = (uint32_t)std::min<uint64_t>(batch, frames - f0);
    EncodeLayout lay;
    size_t bytes = encode_scratch_bytes(nb, frameSize, lastLen, level, &lay);
    void* s = ensure(scratch, bytes);
    if (!s) { r.cudaFailed = true; return r; }
    launches_ += launch_encode_frames(dIn, f0 * frameSize, n, frameSize, nb, level, checksum, s, lay, st);
    if (produced > outCap) { r.zra = 6; return r; }
    launches_ += launch_scan_pack(s, lay, nb, static_cast<uint8_t*>(dOut) + produced, outCap - produced, dSizes64, dTotal, st);
    uint64_t total = 0;
    if (check(cudaMemcpyAsync(&total, dTotal, 8, cudaMemcpyDeviceToHost, st), "size readback") ||
        check(cudaMemcpyAsync(sizesHost + f0, dSizes64, 8 * (size_t)nb, cudaMemcpyDeviceToHost, st), "size readback") ||
        check(cudaStreamSynchronize(st), "encode kernels")) { r.cudaFailed = true; return r; }
    if (produced + total > outCap) { r.zra = 6; return r; }
    produced += total;
  }
  r.total = produced;
  return r;
}

GpuContext::CompressStatus GpuContext::compress_archive(const void* dIn, size_t n, void* dOut, size_t outCap, int level,
                                                        uint32_t frameSize, bool checksum, const uint8_t* metaHost, size_t metaSize,
                                                        bool refMetaQuirk, cudaStream_t st) {
  CompressStatus r;
  bind();
  const uint32_t table = table_entries(n, frameSize);
  const uint64_t frames = table - 1;
  const size_t storedMeta = refMetaQuirk ? 0 : metaSize;
  const size_t tableOff = kFixedHeaderSize + storedMeta;
  const size_t framesOff = tableOff + kEntrySize * (size_t)table;
  if (outCap < framesOff) { r.zra = 6; return r; }
  uint8_t* out = static_cast<uint8_t*>(dOut);
  uint8_t fixed[kFixedHeaderSize];
  write_fixed_header(fixed, n, table, frameSize, (uint32_t)metaSize);
  if (check(cudaMemcpyAsync(out, fixed, sizeof(fixed), cudaMemcpyHostToDevice, st), "header upload")) { r.cudaFailed = true; return r; }
  if (storedMeta && check(cudaMemcpyAsync(out + kFixedHeaderSize, metaHost, storedMeta, cudaMemcpyHostToDevice, st), "meta upload")) {
    r.cudaFailed = true;
    return r;
  }
  const uint32_t lastLen = frames ? (uint32_t)(n - (frames - 1) * (uint64_t)frameSize) : 0;
  uint64_t total = 0;  // compressed bytes so far
  size_t crcLen = metaSize + kEntrySize * (size_t)table;  // what the reference hashes after the fixed part
  uint64_t* dTotal = static_cast<uint64_t*>(ensure(misc, 4096 + crc32_workspace_bytes(crcLen)));
  if (!dTotal) { r.cudaFailed = true; return r; }
  if (check(cudaMemsetAsync(dTotal, 0, 64, st), "memset")) { r.cudaFailed = true; return r; }
  if (frames) {
    EncodeLayout one;
    size_t perFrame = encode_scratch_bytes(1, frameSize, lastLen, level, &one);
    uint64_t batch = std::min<uint64_t>(std::max<uint64_t>(1, enc_budget() / perFrame), frames);
    for (uint64_t f0 = 0; f0 < frames; f0 += batch) {
      uint32_t nb = (uint32_t)std::min<uint64_t>(batch, frames - f0);
      EncodeLayout lay;
      size_t bytes = encode_scratch_bytes(nb, frameSize, lastLen, level, &lay);
      void* s = ensure(scratch, bytes);
      if (!s) { r.cudaFailed = true; return r; }
      launches_ += launch_encode_frames(dIn, f0 * frameSize, n, frameSize, nb, level, checksum, s, lay, st);
      launches_ += launch_scan_gather(s, lay, nb, total, out + tableOff, f0, out + framesOff, outCap - framesOff, dTotal, st);
      if (check(cudaMemcpyAsync(&total, dTotal, 8, cudaMemcpyDeviceToHost, st), "size readback") ||
          check(cudaStreamSynchronize(st), "encode kernels")) { r.cudaFailed = true; return r; }
      if (framesOff + total > outCap) { r.zra = 6; return r; }
    }
  } else {
    // no frames: the table is the single zero sentinel
    if (check(cudaMemsetAsync(out + tableOff, 0, kEntrySize, st), "memset")) { r.cudaFailed = true; return r; }
  }
  if (framesOff + total >= kMaxCompressedSize) { r.zra = 7; return r; }
  // header hash: small fixed part on the host, the (potentially multi-megabyte) table section on the device
  uint32_t* dCrc = reinterpret_cast<uint32_t*>(dTotal) + 4;
  uint32_t crcBig = 0;
  size_t avail = framesOff + total - kFixedHeaderSize;
  size_t hashed = std::min(crcLen, avail);  // the quirk layout may claim more bytes than exist for tiny inputs
  launches_ += launch_crc32(out + kFixedHeaderSize, hashed, dCrc, reinterpret_cast<uint8_t*>(dTotal) + 4096, st);
  if (check(cudaMemcpyAsync(&crcBig, dCrc, 4, cudaMemcpyDeviceToHost, st), "crc readback") ||
      check(cudaStreamSynchronize(st), "crc kernels")) { r.cudaFailed = true; return r; }
  uint32_t crc = crc32_host(fixed, 14);
  crc = crc32_host(fixed + 18, 20, crc);
  crc = crc32_combine(crc, crcBig, hashed);
  uint8_t h[4];
  put_le(h, crc, 4);
  if (check(cudaMemcpyAsync(out + 14, h, 4, cudaMemcpyHostToDevice, st), "hash upload") ||
      check(cudaStreamSynchronize(st), "hash upload")) { r.cudaFailed = true; return r; }
  r.total = framesOff + total;
  return r;
}

}  // namespace zrab
