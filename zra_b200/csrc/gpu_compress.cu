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
                                                        bool refMetaQuirk, cudaStream_t st, const uint8_t* hostIn, uint8_t* hostOut,
                                                        uint8_t* hostFrames) {
  CompressStatus r;
  bind();
  const bool hostIo = (hostOut != nullptr || hostFrames != nullptr) && (hostIn != nullptr || n == 0);
  const uint32_t table = table_entries(n, frameSize);
  const uint64_t frames = table - 1;
  const size_t storedMeta = refMetaQuirk ? 0 : metaSize;
  const size_t tableOff = kFixedHeaderSize + storedMeta;
  const size_t framesOff = tableOff + kEntrySize * (size_t)table;
  if (outCap < framesOff) { r.zra = 6; return r; }
  uint8_t* out = static_cast<uint8_t*>(dOut);
  if (hostIo && !hostFrames) hostFrames = hostOut + framesOff;
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
    if (hostIo) {
      // host pointers: about four batches (measured best of 2..32), at least 32 MiB each (a batch's thread-per-frame
      // stages cost the same 2-3 ms for 256 frames as for 4 096: 64 MiB calls of the streaming Compressor run 7.2 GB/s
      // with 16 MiB batches, 10.9 with 32 MiB), so that the link and the encoder work side by side
      static const uint64_t parts = [] { const char* e = getenv("ZRA_B200_ENC_IO_PARTS"); return e ? std::max<uint64_t>(1, strtoull(e, nullptr, 10)) : 4ull; }();
      static const uint64_t minMb = [] { const char* e = getenv("ZRA_B200_ENC_IO_MIN_MB"); return e ? std::max<uint64_t>(1, strtoull(e, nullptr, 10)) : 32ull; }();
      const uint64_t minFrames = std::max<uint64_t>(1, (minMb << 20) / frameSize);
      batch = std::min(batch, std::max(minFrames, (frames + parts - 1) / parts));
      batch = std::max<uint64_t>(batch, (frames + kMaxChunks - 1) / kMaxChunks);  // one pinned end offset per batch
      if (!ensure_events((frames + batch - 1) / batch + 1)) { r.cudaFailed = true; return r; }
    }
    uint8_t* dInW = static_cast<uint8_t*>(const_cast<void*>(dIn));
    auto upload = [&](uint64_t f0, size_t idx) -> bool {  // the input of the batch that starts at frame f0, on the upload stream
      if (f0 >= frames) return true;
      const uint64_t lo = f0 * frameSize, hi = std::min<uint64_t>(n, (f0 + batch) * (uint64_t)frameSize);
      return !check(cudaMemcpyAsync(dInW + lo, hostIn + lo, hi - lo, cudaMemcpyHostToDevice, upStream_), "input upload") &&
             !check(cudaEventRecord(upEvents_[idx], upStream_), "upload event");
    };
    if (hostIo) {
      if (check(cudaEventRecord(forkEvent_, st), "fork event") || check(cudaStreamWaitEvent(upStream_, forkEvent_, 0), "fork wait") ||
          check(cudaStreamWaitEvent(downStream_, forkEvent_, 0), "fork wait") || !upload(0, 0)) { r.cudaFailed = true; return r; }
    }
    if (!hostIo) {
      for (uint64_t f0 = 0; f0 < frames; f0 += batch) {
        uint32_t nb = (uint32_t)std::min<uint64_t>(batch, frames - f0);
        EncodeLayout lay;
        size_t bytes = encode_scratch_bytes(nb, frameSize, lastLen, level, &lay);
        void* s = ensure(scratch, bytes);
        if (!s) { r.cudaFailed = true; return r; }
        launches_ += launch_encode_frames(dIn, f0 * frameSize, n, frameSize, nb, level, checksum, s, lay, st);
        launches_ += launch_scan_gather(s, lay, nb, kScanContinue, out + tableOff, f0, out + framesOff, outCap - framesOff, dTotal, st);
        if (check(cudaMemcpyAsync(&total, dTotal, 8, cudaMemcpyDeviceToHost, st), "size readback") ||
            check(cudaStreamSynchronize(st), "encode kernels")) { r.cudaFailed = true; return r; }
        if (framesOff + total > outCap) { r.zra = 6; return r; }
      }
    } else {
      // The kernels of every batch are queued without waiting for the host: the running total stays on the device
      // (kScanContinue) and each batch leaves its end offset in pinned memory. The host follows one batch behind and
      // sends batch b's frames down while batch b + 1 is being compressed and batch b + 2 is coming up.
      uint64_t* ends = reinterpret_cast<uint64_t*>(summaryHost_);
      auto retire = [&](size_t b) -> bool {  // batch b has finished: download its frames
        if (check(cudaEventSynchronize(doneEvents_[b]), "encode kernels")) { r.cudaFailed = true; return false; }
        const uint64_t before = total;
        total = ends[b];
        if (framesOff + total > outCap) { r.zra = 6; return false; }
        if (total > before &&
            check(cudaMemcpyAsync(hostFrames + before, out + framesOff + before, total - before, cudaMemcpyDeviceToHost, downStream_),
                  "archive download")) { r.cudaFailed = true; return false; }
        return true;
      };
      auto drain = [&] { cudaStreamSynchronize(pool_[0]); cudaStreamSynchronize(pool_[1]); cudaStreamSynchronize(upStream_); cudaStreamSynchronize(downStream_); };
      // two lanes with a scratch area each: the thread-per-frame stages of a batch (planning, table building) are
      // latency bound and take the same time for 1 000 frames as for 16 000, so they run beside the next batch's matcher
      cudaStream_t lanes[2] = {pool_[0], pool_[1]};
      EncodeLayout layMax;
      const size_t laneBytes = (encode_scratch_bytes((uint32_t)batch, frameSize, lastLen, level, &layMax) + 255) & ~size_t(255);
      uint8_t* s0 = static_cast<uint8_t*>(ensure(scratch, 2 * laneBytes));
      if (!s0) { r.cudaFailed = true; return r; }
      if (check(cudaStreamWaitEvent(lanes[0], forkEvent_, 0), "fork wait") || check(cudaStreamWaitEvent(lanes[1], forkEvent_, 0), "fork wait")) {
        r.cudaFailed = true;
        return r;
      }
      size_t bi = 0;
      for (uint64_t f0 = 0; f0 < frames; f0 += batch, bi++) {
        uint32_t nb = (uint32_t)std::min<uint64_t>(batch, frames - f0);
        cudaStream_t ln = lanes[bi & 1];
        void* s = s0 + (bi & 1) * laneBytes;
        EncodeLayout lay;
        encode_scratch_bytes(nb, frameSize, lastLen, level, &lay);
        if (!upload(f0 + batch, bi + 1) || check(cudaStreamWaitEvent(ln, upEvents_[bi], 0), "upload wait")) { drain(); r.cudaFailed = true; return r; }
        launches_ += launch_encode_frames(dIn, f0 * frameSize, n, frameSize, nb, level, checksum, s, lay, ln);
        // the running total is handed from batch to batch
        if (bi > 0 && check(cudaStreamWaitEvent(ln, doneEvents_[bi - 1], 0), "batch order")) { drain(); r.cudaFailed = true; return r; }
        launches_ += launch_scan_gather(s, lay, nb, kScanContinue, out + tableOff, f0, out + framesOff, outCap - framesOff, dTotal, ln);
        if (check(cudaMemcpyAsync(&ends[bi], dTotal, 8, cudaMemcpyDeviceToHost, ln), "size readback") ||
            check(cudaEventRecord(doneEvents_[bi], ln), "batch event")) { drain(); r.cudaFailed = true; return r; }
        if (bi > 0 && !retire(bi - 1)) { drain(); return r; }
      }
      if (!retire(bi - 1)) { drain(); return r; }
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
  if (hostIo) {  // header, metadata, seek table; then every download must have landed
    if ((hostOut && (check(cudaMemcpyAsync(hostOut, out, std::min<size_t>(framesOff, r.total), cudaMemcpyDeviceToHost, st), "header download") ||
                     check(cudaStreamSynchronize(st), "header download"))) ||
        check(cudaStreamSynchronize(downStream_), "archive download")) {
      r.cudaFailed = true;
      return r;
    }
  }
  return r;
}

}  // namespace zrab
