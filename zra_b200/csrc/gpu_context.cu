// gpu_context.cu — GPU context: device binding, buffers, and the batched decode driver.
#include "gpu_context.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

namespace zrab {

GpuContext::GpuContext(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    lastError_ = std::string("zra-b200: no usable CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path";
    cudaGetLastError();
    return;
  }
  if (device < 0) {
    if (check(cudaGetDevice(&device), "cudaGetDevice")) return;
  }
  device_ = device;
  if (check(cudaSetDevice(device_), "cudaSetDevice")) return;
  if (check(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate")) return;
  ok_ = true;
}

GpuContext::~GpuContext() {
  if (!ok_) return;
  cudaSetDevice(device_);
  for (DevBuf* b : {&scratch, &stageIn, &stageOut, &misc})
    if (b->p) cudaFree(b->p);
  if (stream_) cudaStreamDestroy(stream_);
}

void GpuContext::bind() { cudaSetDevice(device_); }

bool GpuContext::check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return false;
  lastError_ = std::string("zra-b200: CUDA failure in ") + what + ": " + cudaGetErrorString(e);
  cudaGetLastError();
  return true;
}

void* GpuContext::ensure(DevBuf& b, size_t bytes) {
  if (bytes <= b.cap && b.p) return b.p;
  if (b.p) {
    cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = std::max<size_t>(bytes + bytes / 8, 1 << 20);
  want = (want + 255) & ~size_t(255);
  if (check(cudaMalloc(&b.p, want), "cudaMalloc")) {
    // retry with the exact size before giving up
    want = (bytes + 255) & ~size_t(255);
    if (check(cudaMalloc(&b.p, want), "cudaMalloc")) {
      b.p = nullptr;
      return nullptr;
    }
  }
  b.cap = want;
  return b.p;
}

static size_t scratch_budget() {
  static size_t v = [] {
    const char* s = getenv("ZRA_B200_SCRATCH_MB");
    size_t mb = s ? strtoull(s, nullptr, 10) : 6144;
    return std::max<size_t>(mb, 64) << 20;
  }();
  return v;
}

DecodeResult GpuContext::decode(const void* dSrc, size_t srcSize, const HostFrame* frames, const ArchiveInfo* info,
                                uint64_t firstFrame, uint64_t nFrames, uint32_t maxDstCap, void* dDst, uint32_t* frameSizes,
                                cudaStream_t st) {
  DecodeResult res;
  (void)srcSize;
  if (!nFrames) return res;
  bind();
  // batch size: as many frames as the scratch budget allows
  DecodeLayout one;
  size_t perFrame = decode_scratch_bytes(1, maxDstCap, &one);
  uint64_t batch = std::max<uint64_t>(1, scratch_budget() / perFrame);
  batch = std::min<uint64_t>(batch, nFrames);
  batch = std::min<uint64_t>(batch, 1u << 22);
  const uint32_t baseRounds = std::max<uint32_t>(1, (maxDstCap + (1u << 17) - 1) >> 17);
  std::vector<uint8_t> ctxHost;

  for (uint64_t f0 = 0; f0 < nFrames; f0 += batch) {
    uint32_t n = (uint32_t)std::min<uint64_t>(batch, nFrames - f0);
    DecodeLayout lay;
    size_t bytes = decode_scratch_bytes(n, maxDstCap, &lay);
    uint8_t* s = static_cast<uint8_t*>(ensure(scratch, bytes));
    if (!s) { res.cudaFailed = true; return res; }
    launch_summary_reset(s, lay, st);
    uint64_t dstBase = 0;
    if (frames) {
      if (check(cudaMemcpyAsync(s + lay.offDescs, frames + f0, sizeof(HostFrame) * (size_t)n, cudaMemcpyHostToDevice, st),
                "descriptor upload")) { res.cudaFailed = true; return res; }
    } else {
      dstBase = firstFrame * info->frameSize;
      launch_build_descs(dSrc, 38ull + info->metaSize, info->headerSize, srcSize, info->uncompressedSize, info->frameSize,
                         (uint32_t)(firstFrame + f0), n, dstBase, s, lay, st);
      launches_ += 1;
    }
    uint32_t rounds = baseRounds;
    bool first = true;
    uint32_t summary[4];
    for (int pass = 0;; pass++) {
      launch_decode_rounds(dSrc, dDst, n, rounds, first, s, lay, st);
      launch_frame_finish(dSrc, dDst, n, s, lay, st);
      launches_ += 4ull * rounds + 1;
      first = false;
      if (check(cudaMemcpyAsync(summary, s + lay.offSummary, sizeof(summary), cudaMemcpyDeviceToHost, st), "summary readback") ||
          check(cudaStreamSynchronize(st), "decode kernels")) { res.cudaFailed = true; return res; }
      if (summary[0] != 0xFFFFFFFFu || summary[1] == 0) break;
      // frames with more blocks than the zstd encoder would emit: keep going
      rounds = std::min<uint32_t>(rounds * 2, 64);
      if (pass > 1 << 16) break;
    }
    if (summary[0] != 0xFFFFFFFFu) {
      uint32_t code = 0;
      size_t off = lay.offCtxs + (size_t)summary[0] * frame_ctx_size() + frame_status_offset();
      if (check(cudaMemcpy(&code, s + off, 4, cudaMemcpyDeviceToHost), "status readback")) { res.cudaFailed = true; return res; }
      res.zstd = (int)code;
      res.failedFrame = (uint32_t)(f0 + summary[0]);
      return res;
    }
    if (frameSizes) {
      ctxHost.resize((size_t)n * frame_ctx_size());
      if (check(cudaMemcpy(ctxHost.data(), s + lay.offCtxs, ctxHost.size(), cudaMemcpyDeviceToHost), "ctx readback")) { res.cudaFailed = true; return res; }
      for (uint32_t i = 0; i < n; i++) {
        uint32_t v;
        memcpy(&v, ctxHost.data() + (size_t)i * frame_ctx_size() + 12, 4);  // FrameCtx::dstPos
        frameSizes[f0 + i] = v;
      }
    }
  }
  return res;
}

GpuContext* default_context() {
  thread_local std::unique_ptr<GpuContext> ctx;
  if (!ctx) ctx.reset(new GpuContext(-1));
  return ctx.get();
}

}  // namespace zrab
