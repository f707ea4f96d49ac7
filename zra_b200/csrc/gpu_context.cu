// gpu_context.cu — GPU context: device binding, buffers, and the chunked multi-stream decode driver.
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
  {
    // chunk i runs on pool_[i]: earlier chunks get the higher priority, so that the tail kernels of an early chunk
    // (whose output the host is waiting to download) are not queued behind a later chunk's whole-SM CTAs
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    static const bool flat = getenv("ZRA_B200_FLAT_PRIORITY") != nullptr;
    for (int i = 0; i < kPoolStreams; i++) {
      int prio = flat ? least : std::min(least, greatest + i);
      if (check(cudaStreamCreateWithPriority(&pool_[i], cudaStreamNonBlocking, prio), "cudaStreamCreate")) return;
      if (check(cudaStreamCreateWithPriority(&side_[i].st, cudaStreamNonBlocking, prio), "cudaStreamCreate")) return;
      if (check(cudaEventCreateWithFlags(&side_[i].fork, cudaEventDisableTiming), "cudaEventCreate")) return;
      if (check(cudaEventCreateWithFlags(&side_[i].join, cudaEventDisableTiming), "cudaEventCreate")) return;
    }
  }
  if (check(cudaEventCreateWithFlags(&forkEvent_, cudaEventDisableTiming), "cudaEventCreate")) return;
  {
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    if (check(cudaStreamCreateWithPriority(&upStream_, cudaStreamNonBlocking, greatest), "cudaStreamCreate")) return;
    if (check(cudaStreamCreateWithPriority(&downStream_, cudaStreamNonBlocking, greatest), "cudaStreamCreate")) return;
  }
  if (check(cudaMallocHost(&summaryHost_, sizeof(uint32_t) * 4 * kMaxChunks), "cudaMallocHost")) return;
  ok_ = true;
}

GpuContext::~GpuContext() {
  if (!ok_) return;
  cudaSetDevice(device_);
  if (raHost_) cudaFreeHost(raHost_);
  for (DevBuf* b : {&scratch, &stageIn, &stageOut, &misc, &raSlotOf, &raUnique, &raDescs, &raFrames})
    if (b->p) cudaFree(b->p);
  for (auto& s : pool_)
    if (s) cudaStreamDestroy(s);
  for (auto& l : side_) {
    if (l.st) cudaStreamDestroy(l.st);
    if (l.fork) cudaEventDestroy(l.fork);
    if (l.join) cudaEventDestroy(l.join);
  }
  if (forkEvent_) cudaEventDestroy(forkEvent_);
  for (cudaEvent_t e : upEvents_) cudaEventDestroy(e);
  for (cudaEvent_t e : doneEvents_) cudaEventDestroy(e);
  if (upStream_) cudaStreamDestroy(upStream_);
  if (downStream_) cudaStreamDestroy(downStream_);
  if (summaryHost_) cudaFreeHost(summaryHost_);
  for (uint8_t* p : pinnedStage_) if (p) cudaFreeHost(p);
  if (stream_) cudaStreamDestroy(stream_);
}

void GpuContext::bind() { cudaSetDevice(device_); }

bool GpuContext::ensure_events(size_t n) {
  while (upEvents_.size() < n) {
    cudaEvent_t a = nullptr, b = nullptr;
    if (check(cudaEventCreateWithFlags(&a, cudaEventDisableTiming), "cudaEventCreate")) return false;
    upEvents_.push_back(a);
    if (check(cudaEventCreateWithFlags(&b, cudaEventDisableTiming), "cudaEventCreate")) return false;
    doneEvents_.push_back(b);
  }
  return true;
}

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

uint8_t* GpuContext::pinned_stage(size_t bytes, int which) {
  which &= 1;
  stageGen_[which]++;
  if (bytes <= pinnedStageCap_[which] && pinnedStage_[which]) return pinnedStage_[which];
  bind();
  if (pinnedStage_[which]) {
    cudaFreeHost(pinnedStage_[which]);
    pinnedStage_[which] = nullptr;
    pinnedStageCap_[which] = 0;
  }
  size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 20);
  void* p = nullptr;
  if (check(cudaMallocHost(&p, want), "cudaMallocHost")) {
    want = std::max<size_t>(bytes, 1);
    if (check(cudaMallocHost(&p, want), "cudaMallocHost")) return nullptr;
  }
  pinnedStage_[which] = static_cast<uint8_t*>(p);
  pinnedStageCap_[which] = want;
  return pinnedStage_[which];
}

static size_t scratch_budget() {
  static size_t v = [] {
    const char* s = getenv("ZRA_B200_SCRATCH_MB");
    size_t mb = s ? strtoull(s, nullptr, 10) : 8192;
    return std::max<size_t>(mb, 64) << 20;
  }();
  return v;
}

static uint32_t chunk_target() {
  static uint32_t v = [] {
    const char* s = getenv("ZRA_B200_CHUNKS");
    // device-resident calls: 2 chunks (1 GiB of 64 KiB frames: 1 -> 5.88, 2 -> 5.19, 4 -> 5.6 .. 6.1 ms; gpurun_out/r03l)
    uint32_t n = s ? (uint32_t)strtoul(s, nullptr, 10) : 2;
    return std::max<uint32_t>(1, std::min<uint32_t>(n, 256));
  }();
  return v;
}

// Host-pointer calls (uploads and downloads inside the chunk pipeline) are PCIe-bound. Finer or geometrically
// growing chunks were measured SLOWER (profiles/r01e, r01h): every chunk's kernels take about the same time whatever
// its size (serial per-frame chains), and more chunks in flight only delay each other's tails.
static uint32_t io_chunk_target() {
  static uint32_t v = [] {
    const char* s = getenv("ZRA_B200_IO_CHUNKS");
    uint32_t n = s ? (uint32_t)strtoul(s, nullptr, 10) : 4;  // measured (profiles/r01e): 4 -> 42.7 GB/s, 8 -> 42.4, 16 -> 38.4, 32 -> 32.9
    return std::max<uint32_t>(1, std::min<uint32_t>(n, 256));
  }();
  return v;
}

namespace {
  struct Chunk {
    uint64_t f0;
    uint32_t n;
    DecodeLayout lay;
    uint8_t* scratch;
    cudaStream_t st;
    uint32_t idx;
    uint32_t* summary;  // pinned host words
  };
}  // namespace

DecodeResult GpuContext::decode(const void* dSrc, size_t srcSize, const HostFrame* frames, const ArchiveInfo* info,
                                uint64_t firstFrame, uint64_t nFrames, uint32_t maxDstCap, void* dDst, uint32_t* frameSizes,
                                cudaStream_t st, const HostStaging* io) {
  DecodeResult res;
  if (!nFrames) return res;
  bind();
  DecodeLayout one;
  const size_t perFrame = decode_scratch_bytes(1, maxDstCap, &one);
  // frames per group = what the scratch budget holds; a group is cut into chunks that run concurrently
  const uint64_t groupFrames = std::min<uint64_t>(std::max<uint64_t>(1, scratch_budget() / perFrame), 1u << 22);
  const uint32_t baseRounds = std::max<uint32_t>(1, (maxDstCap + (1u << 17) - 1) >> 17);
  static const bool timeline = getenv("ZRA_B200_TIMELINE") != nullptr;  // tuning aid: keep the chunk pipeline, dump every mark
  const bool single = profiling_ && !timeline;  // per-kernel event timing needs one chunk on one stream
  KernelTimer* tm = profiling_ ? &timer : nullptr;
  std::vector<uint8_t> ctxHost;
  std::vector<Chunk> chunks;
  auto fail_cuda = [&]() { res.cudaFailed = true; return res; };

  for (uint64_t g0 = 0; g0 < nFrames; g0 += groupFrames) {
    const uint64_t gN = std::min<uint64_t>(groupFrames, nFrames - g0);
    // ---- plan the chunks of this group
    const bool hostIo = io && (io->hostSrc || io->hostDst);
    uint32_t nChunks = single ? 1u : (uint32_t)std::min<uint64_t>(hostIo ? io_chunk_target() : chunk_target(), std::max<uint64_t>(1, gN / 64));
    nChunks = std::min<uint32_t>(nChunks, kMaxChunks);
    const uint64_t per = (gN + nChunks - 1) / nChunks;
    chunks.clear();
    size_t total = 0;
    auto add_chunk = [&](uint64_t c0, uint64_t n) {
      Chunk c;
      c.f0 = g0 + c0;
      c.n = (uint32_t)n;
      total += decode_scratch_bytes(c.n, maxDstCap, &c.lay);
      chunks.push_back(c);
    };
    // tuning aids: explicit chunk sizes "a,b,c" (the rest = one more chunk) for device-resident / host-pointer calls
    static const char* planEnv = getenv("ZRA_B200_PLAN");
    static const char* ioPlanEnv = getenv("ZRA_B200_IO_PLAN");
    const char* plan = hostIo ? ioPlanEnv : planEnv;
    if (plan && !single) {
      uint64_t c0 = 0;
      for (const char* q = plan; *q && c0 < gN;) {
        uint64_t n = std::min<uint64_t>(strtoull(q, const_cast<char**>(&q), 10), gN - c0);
        if (n) add_chunk(c0, n);
        c0 += n;
        if (*q == ',') q++; else break;
      }
      if (c0 < gN) add_chunk(c0, gN - c0);
    } else {
      for (uint64_t c0 = 0; c0 < gN; c0 += per) add_chunk(c0, std::min<uint64_t>(per, gN - c0));
    }
    uint8_t* base = static_cast<uint8_t*>(ensure(scratch, total));
    if (!base) return fail_cuda();
    size_t off = 0;
    for (size_t i = 0; i < chunks.size(); i++) {
      DecodeLayout tmp;
      chunks[i].scratch = base + off;
      off += decode_scratch_bytes(chunks[i].n, maxDstCap, &tmp);
      chunks[i].st = single ? st : pool_[i % kPoolStreams];
      chunks[i].idx = (uint32_t)i;
      chunks[i].summary = summaryHost_ + 4 * i;
    }
    // ---- fork: the pool streams start after whatever the caller queued on `st`
    if (!single) {
      if (check(cudaEventRecord(forkEvent_, st), "fork event")) return fail_cuda();
      for (int i = 0; i < kPoolStreams && i < (int)chunks.size(); i++)
        if (check(cudaStreamWaitEvent(pool_[i], forkEvent_, 0), "fork wait")) return fail_cuda();
      if (hostIo && (check(cudaStreamWaitEvent(upStream_, forkEvent_, 0), "fork wait") || !ensure_events(chunks.size()))) return fail_cuda();
    }
    // ---- enqueue every chunk: [H2D] -> descriptors -> rounds -> finish -> summary [-> D2H]
    // Copies normally ride on the chunks' own streams (the copy engines take them in issue order). Putting all uploads
    // and all downloads on two dedicated streams in chunk order was measured equal at 4 chunks and slower beyond
    // (profiles/r01q); it stays selectable for experiments. The host path is link-bound: 1 GiB out + 0.33 GiB in
    // take 24.2 ms (44.4 GB/s) against 56 GB/s one-way / 49.7 GB/s duplex measured on the same box (profiles/r01r).
    static const bool orderedEnv = getenv("ZRA_B200_ORDERED_IO") != nullptr;
    const bool ordered = hostIo && !single && orderedEnv;
    auto enqueue_tail = [&](Chunk& c) -> bool {
      launch_frame_finish(dSrc, dDst, c.n, c.scratch, c.lay, c.st, tm);
      launches_ += 1;
      if (check(cudaMemcpyAsync(c.summary, c.scratch + c.lay.offSummary, 16, cudaMemcpyDeviceToHost, c.st), "summary readback"))
        return false;
      if (ordered && check(cudaEventRecord(doneEvents_[c.idx], c.st), "chunk done event")) return false;
      return true;
    };
    // output download of one chunk (second pass: with pageable host memory cudaMemcpyAsync blocks the
    // host until the chunk is done, which must not delay the enqueueing of the other chunks)
    auto enqueue_download = [&](Chunk& c) -> bool {
      if (!(io && io->hostDst && frames)) return true;
      const HostFrame& a = frames[c.f0];
      const HostFrame& b = frames[c.f0 + c.n - 1];
      uint64_t lo = std::max<uint64_t>(a.dstOff, io->dstSkip);
      uint64_t hi = std::min<uint64_t>(b.dstOff + b.dstCap, io->dstSize == ~0ull ? ~0ull : io->dstSkip + io->dstSize);
      if (hi <= lo) return true;
      if (ordered && check(cudaStreamWaitEvent(downStream_, doneEvents_[c.idx], 0), "download wait")) return false;
      return !check(cudaMemcpyAsync(io->hostDst + (lo - io->dstSkip), static_cast<const uint8_t*>(dDst) + lo, hi - lo,
                                    cudaMemcpyDeviceToHost, ordered ? downStream_ : c.st), "output download");
    };
    for (Chunk& c : chunks) {
      launch_summary_reset(c.scratch, c.lay, c.st);
      if (frames) {
        // `frames` may also be a DEVICE array (batched random access builds its descriptors on the GPU). The
        // descriptors go first: from pageable host memory this copy blocks the host until everything queued on the
        // stream before it is done, which must not be the chunk's source upload.
        if (check(cudaMemcpyAsync(c.scratch + c.lay.offDescs, frames + c.f0, sizeof(HostFrame) * (size_t)c.n, cudaMemcpyDefault,
                                  c.st), "descriptor upload"))
          return fail_cuda();
        if (io && io->hostSrc) {
          const HostFrame& a = frames[c.f0];
          const HostFrame& b = frames[c.f0 + c.n - 1];
          uint64_t lo = a.srcOff, hi = b.srcOff + b.srcLen;
          if (hi > lo && check(cudaMemcpyAsync(const_cast<uint8_t*>(static_cast<const uint8_t*>(dSrc)) + lo, io->hostSrc + lo, hi - lo,
                                               cudaMemcpyHostToDevice, ordered ? upStream_ : c.st), "source upload"))
            return fail_cuda();
          if (ordered && (check(cudaEventRecord(upEvents_[c.idx], upStream_), "upload event") ||
                          check(cudaStreamWaitEvent(c.st, upEvents_[c.idx], 0), "upload wait")))
            return fail_cuda();
        }
      } else {
        launch_build_descs(dSrc, 38ull + info->metaSize, info->headerSize, srcSize, info->uncompressedSize, info->frameSize,
                           (uint32_t)(firstFrame + c.f0), c.n, firstFrame * info->frameSize, c.scratch, c.lay, c.st, tm);
        launches_ += 1;
      }
      static const bool sideHuf = [] { const char* e = getenv("ZRA_B200_SIDE_HUF"); return !e || atoi(e) != 0; }();
      const SideLane* side = (single || !sideHuf || chunks.size() > (size_t)kPoolStreams) ? nullptr : &side_[c.idx % kPoolStreams];
      launches_ += launch_decode_rounds(dSrc, dDst, c.n, baseRounds, true, c.scratch, c.lay, c.st, tm, side);
      if (!enqueue_tail(c)) return fail_cuda();
    }
    for (Chunk& c : chunks)
      if (!enqueue_download(c)) return fail_cuda();
    if (io && io->whileBusy && g0 == 0) (*io->whileBusy)();  // everything of this group is queued: the GPU works, the caller reads ahead
    // ---- join, then look at every chunk's summary (in frame order, so the lowest failing frame wins)
    if (tm && timeline) {
      for (Chunk& c : chunks) cudaStreamSynchronize(c.st);
      tm->dump_timeline(stderr);
    }
    for (Chunk& c : chunks) {
      if (check(cudaStreamSynchronize(c.st), "decode kernels")) return fail_cuda();
      if (tm) tm->collect();
      uint32_t rounds = baseRounds;
      bool extra = false;
      for (int pass = 0; c.summary[0] == 0xFFFFFFFFu && c.summary[1] != 0 && pass < (1 << 16); pass++) {
        // frames with more blocks than the zstd encoder would emit: keep going (rare, serial)
        rounds = std::min<uint32_t>(rounds * 2, 64);
        launches_ += launch_decode_rounds(dSrc, dDst, c.n, rounds, false, c.scratch, c.lay, c.st, tm);
        if (!enqueue_tail(c) || check(cudaStreamSynchronize(c.st), "decode kernels")) return fail_cuda();
        if (tm) tm->collect();
        extra = true;
      }
      if (extra && c.summary[0] == 0xFFFFFFFFu &&
          (!enqueue_download(c) || check(cudaStreamSynchronize(ordered ? downStream_ : c.st), "output download")))
        return fail_cuda();
      if (c.summary[0] != 0xFFFFFFFFu) {
        uint32_t code = 0;
        size_t o = c.lay.offCtxs + (size_t)c.summary[0] * frame_ctx_size() + frame_status_offset();
        if (check(cudaMemcpy(&code, c.scratch + o, 4, cudaMemcpyDeviceToHost), "status readback")) return fail_cuda();
        for (Chunk& rest : chunks) cudaStreamSynchronize(rest.st);
        if (ordered) cudaStreamSynchronize(downStream_);
        res.failedFrame = (uint32_t)(c.f0 + c.summary[0]);
        // The reference decodes a range of frames with ZSTD_decompressMultiFrame, which reports an unknown magic
        // number AFTER a frame that decoded as srcSize_wrong ("following bytes are garbage", zstd_decompress.c:773-781):
        // the lowest failing frame's prefix_unknown (10) becomes 72 unless it is the first frame of the call.
        if (code == 10u && res.failedFrame > 0) code = 72u;
        res.zstd = (int)code;
        return res;
      }
      if (frameSizes) {
        ctxHost.resize((size_t)c.n * frame_ctx_size());
        if (check(cudaMemcpy(ctxHost.data(), c.scratch + c.lay.offCtxs, ctxHost.size(), cudaMemcpyDeviceToHost), "ctx readback"))
          return fail_cuda();
        for (uint32_t i = 0; i < c.n; i++) {
          uint32_t v;
          memcpy(&v, ctxHost.data() + (size_t)i * frame_ctx_size() + 12, 4);  // FrameCtx::dstPos
          frameSizes[c.f0 + i] = v;
        }
      }
    }
    if (ordered && check(cudaStreamSynchronize(downStream_), "output download")) return fail_cuda();
  }
  return res;
}

GpuContext* default_context() {
  thread_local std::unique_ptr<GpuContext> ctx;
  if (!ctx) ctx.reset(new GpuContext(-1));
  return ctx.get();
}

}  // namespace zrab
