// decode_kernels.cu — sm_100a kernels of the frame-parallel zstd decoder and their launcher.
//
// Kernel inventory (one "round" = block r of every frame in the batch; see decode_core.cuh):
//   k_build_descs   1 thread / frame   40-bit seek-table entries -> FrameDesc (ZRA geometry)
//   k_block_setup   1 thread / frame   headers + Huffman/FSE table construction, work lists
//   k_huf_decode    persistent warps   4 lanes per frame (one per Huffman stream), 8 frames per warp,
//                                      decode table staged in shared memory, lane-quads pull frames
//                                      from a work list as they finish
//   k_seq_decode    persistent warps   1 lane per frame, 32 frames per warp advance in lock-step,
//                                      the three compact FSE tables of every frame in shared memory
//                                      (2.5 KiB per frame), lanes pull frames from a work list
//   k_seq_execute   1 warp / frame     literal/match copies into the output, raw/RLE blocks
//   k_frame_finish  4 threads/ frame   XXH64 checksum + final size checks + error summary
// Replaces the serial loop of ZSTD_decompressMultiFrame that zra::DecompressBuffer /
// DecompressRA / Decompressor / FullDecompressor drive (source/zra.cpp:249,280-293,397-410,435).
#include <cuda_runtime.h>

#include <cstdlib>

#include "decode_core.cuh"
#include "decode_exec.cuh"
#include "decode_launch.h"
#include "decode_seq.cuh"
#include "xxh64.cuh"

namespace zrab {

constexpr u32 kFull = 0xFFFFFFFFu;
constexpr u32 kNone = 0xFFFFFFFFu;

// ------------------------------------------------------------------------------------------
// Seek table (5-byte little-endian entries, offsets relative to the end of the header) -> descs.
// Frame i occupies [hdr + E[i], hdr + E[i+1]) and regenerates min(frameSize, U - i*frameSize).
__global__ void k_build_descs(const u8* __restrict__ archive, u64 tableOff, u64 headerSize, u64 archiveSize,
                              u64 uncompressedSize, u32 frameSize, u32 firstFrame, u32 nFrames, u64 dstBase,
                              FrameDesc* __restrict__ descs, u32* __restrict__ summary) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nFrames) return;
  u64 f = (u64)firstFrame + i;
  const u8* e = archive + tableOff + 5 * f;
  u64 a = (u64)ld32(e) | ((u64)e[4] << 32);
  u64 b = (u64)ld32(e + 5) | ((u64)e[9] << 32);
  FrameDesc d;
  u64 begin = f * frameSize;
  u64 left = uncompressedSize - begin;
  d.srcOff = headerSize + a;
  d.dstOff = begin - dstBase;
  d.dstCap = (u32)(left < frameSize ? left : frameSize);
  d.exact = 1;
  d.pad = 0;
  if (b < a || headerSize + b > archiveSize || b - a > 0xFFFFFFFFull) {
    d.srcLen = 0;  // decodes to ZE_SRC_WRONG
    atomicMin(&summary[2], i);
  } else {
    d.srcLen = (u32)(b - a);
  }
  descs[i] = d;
}

// One warp per CTA, one frame per lane. The sequence tables are BUILT IN SHARED MEMORY (the build is a
// read-modify-write of every cell: spread the symbols, then number the cells; in HBM scratch each of those 2 x 1280
// dependent accesses is an L2 round trip), one table at a time (block_setup_head / _table x 3 / _tail): each lane
// has a 1 KiB staging area (one word of padding, so that lanes touching the same cell index fall into different
// banks), and after each table the whole warp copies the rebuilt ones out with coalesced word stores. 32.1 KiB per
// CTA: seven warps per SM, which is what bounds this kernel on archives of many small frames.
constexpr u32 kSetupStageWords = 512 * sizeof(CSym) / 4 + 1;  // 257
constexpr u32 kSetupSmem = 32 * kSetupStageWords * 4;          // 32.1 KiB

__global__ void __launch_bounds__(32) k_block_setup(const u8* __restrict__ src, const FrameDesc* __restrict__ descs,
                                                    FrameCtx* __restrict__ ctxs, FrameTables* __restrict__ tabs, u32 nFrames,
                                                    u32 firstRound, RoundWork* __restrict__ work, u32* __restrict__ hufList,
                                                    u32* __restrict__ seqList, u32 splitSmall) {
  extern __shared__ __align__(16) u8 smem[];
  u32* stageWords = reinterpret_cast<u32*>(smem);
  const u32 lane = threadIdx.x;
  const u32 i = blockIdx.x * 32 + lane;
  const bool have = i < nFrames;
  CSym* stage = reinterpret_cast<CSym*>(stageWords + lane * kSetupStageWords);
  FrameCtx c;
  SetupCursor cur;
  cur.live = false;
  if (have) {
    c = ctxs[i];
    cur = block_setup_head(src, descs[i], c, tabs[i], firstRound != 0);
  }
  // tables in stream order; FrameTables word offsets: ll 0 (256 words), ml 256 (256), of 512 (128)
  for (u32 t = 0; t < 3; t++) {
    const u32 kind = t == 0 ? SEQ_LL : (t == 1 ? SEQ_OF : SEQ_ML);
    const u32 toWord = kind == SEQ_LL ? 0u : (kind == SEQ_ML ? 256u : 512u);
    const u32 words = kind == SEQ_OF ? 128u : 256u;
    bool built = false;
    if (have) built = block_setup_table(src, descs[i], c, cur, kind, stage);
    __syncwarp();
    u32 any = __ballot_sync(kFull, built);
    while (any) {
      const int who = __ffs(any) - 1;
      any &= any - 1;
      const u32* from = stageWords + who * kSetupStageWords;
      u32* to = reinterpret_cast<u32*>(&tabs[blockIdx.x * 32 + who]) + toWord;
      for (u32 k = lane; k < words; k += 32) to[k] = from[k];
    }
    __syncwarp();
  }
  if (!have) return;
  block_setup_tail(descs[i], c, cur);
  ctxs[i] = c;
  if (c.blkType == BT_COMPRESSED && !c.status) {
    if (c.litMode == LIT_HUF && c.litSize) hufList[atomicAdd(&work->hufCount, 1u)] = i;
    if (c.nbSeq) {
      if (splitSmall && c.llLog <= kSeqSmallLogMax && c.mlLog <= kSeqSmallLogMax) seqList[nFrames - 1 - atomicAdd(&work->seqCountS, 1u)] = i;
      else seqList[atomicAdd(&work->seqCount, 1u)] = i;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Huffman literals. One warp = 8 frame slots x 4 lanes (lane & 3 = stream). The quad leader BUILDS
// the frame's decode table from its weights straight into the quad's shared-memory slot, in the
// two-level form (entropy.cuh: typically < 0.8 KiB instead of 4 KiB), so that every frame in flight
// on an SM keeps its table on chip; a tree that does not fit the 1 KiB slot is built in the frame's
// global scratch instead and read through generic loads. The streams are read with the same
// warp-synchronous ring reader as the sequence stream. Symbols are decoded in groups of four (one
// 64-bit window, one aligned 32-bit store); the warp runs a uniform number of groups between checks.
constexpr u32 kHufSlotEntries = 512;
constexpr u32 kHufWarpSmem = 8 * kHufSlotEntries * sizeof(HufSym) + kRingWords * 32 * sizeof(u32);  // 10 KiB

__global__ void __launch_bounds__(32) k_huf_decode(const u8* __restrict__ src, const FrameDesc* __restrict__ descs,
                                                   FrameCtx* __restrict__ ctxs, FrameTables* __restrict__ tabs,
                                                   u8* __restrict__ lit, u32 litStride, RoundWork* __restrict__ work,
                                                   const u32* __restrict__ hufList) {
  extern __shared__ __align__(16) u8 smem[];
  HufSym* slots = reinterpret_cast<HufSym*>(smem);
  u32* ring = reinterpret_cast<u32*>(smem + 8 * kHufSlotEntries * sizeof(HufSym)) + threadIdx.x;  // [kRingWords][32]
  const u32 lane = threadIdx.x, quad = lane >> 2, s = lane & 3;
  const u32 total = work->hufCount;
  const HufSym* table = slots;
  HufLevels lv;
  lv.log = lv.p = lv.nEsc = lv.l2 = 0;
  u32 frame = kNone;      // frame owned by this quad
  bool quadIdle = true;   // quad has no frame
  bool exhausted = false;
  bool laneDone = true;   // this lane's stream is finished (or it has none)
  SeqReader br;
  br.p = br.b0 = br.loadedW = br.reqW = 0;
  u32 i = 0, n = 0;
  u8* out = nullptr;
  for (;;) {
    // ---- hand frames to idle quads
    if (__any_sync(kFull, quadIdle && !exhausted)) {
      u32 f = kNone;
      if (quadIdle && !exhausted && s == 0) {
        u32 k = atomicAdd(&work->hufNext, 1u);
        if (k < total) f = hufList[k];
      }
      f = __shfl_sync(kFull, f, lane & ~3u);
      if (quadIdle && !exhausted && f == kNone) exhausted = true;
      u32 where = 0;  // 1: shared-memory slot, 2: global scratch
      if (f != kNone && s == 0) {
        const FrameCtx& c = ctxs[f];
        HufSym* slot = slots + quad * kHufSlotEntries;
        if (huf_build_two_level(slot, kHufSlotEntries, tabs[f].hufWeights, c.hufCount, c.hufLog, &lv)) where = 1;
        else { huf_build_two_level(tabs[f].huf, kHufGlobalCap, tabs[f].hufWeights, c.hufCount, c.hufLog, &lv); where = 2; }
      }
      __syncwarp();
      {  // the leader's table description goes to its quad (all lanes take part in the shuffles)
        const u32 leader = lane & ~3u;
        const u32 w = __shfl_sync(kFull, where, leader);
        const u32 a = __shfl_sync(kFull, lv.log, leader), b = __shfl_sync(kFull, lv.p, leader);
        const u32 c = __shfl_sync(kFull, lv.nEsc, leader), d = __shfl_sync(kFull, lv.l2, leader);
        if (f != kNone) { where = w; lv.log = a; lv.p = b; lv.nEsc = c; lv.l2 = d; }
      }
      if (f != kNone) {
        frame = f;
        quadIdle = false;
        const FrameCtx& c = ctxs[f];
        table = where == 1 ? slots + quad * kHufSlotEntries : tabs[f].huf;
        laneDone = true;
        if (s < c.nStreams) {
          u32 seg = c.nStreams == 4 ? (c.litSize + 3) / 4 : c.litSize;
          n = (c.nStreams == 4 && s == 3) ? c.litSize - 3 * seg : seg;
          out = lit + (u64)f * litStride + s * seg;
          i = 0;
          laneDone = false;
          if (!br.init(src, descs[f].srcOff + c.strOff[s], c.strLen[s], ring, 32)) {
            ctxs[f].status = (u32)ZE_CORRUPTION;  // literals fail first in the reference: this code wins over the sequence stage's
            laneDone = true;
          } else {
            // head: bring the output cursor to a 4-byte boundary so that the groups store whole words
            u32 head = (4u - (u32)(reinterpret_cast<uintptr_t>(out) & 3u)) & 3u;
            if (head > n) head = n;
            for (; i < head; i++) {
              u32 hi, lo;
              br.window(hi, lo);
              const u32 e = huf_lookup(table, lv, hi);
              out[i] = (u8)e;
              br.p -= (i32)(e >> 8);
            }
          }
        }
      }
    }
    if (__all_sync(kFull, quadIdle)) break;
    // ---- a warp-uniform number of 4-symbol groups
    const u32 steps = __reduce_min_sync(kFull, laneDone ? 0xFFFFFFFFu : (n - i) >> 2);
    if (!laneDone && steps != 0xFFFFFFFFu) {
      u32* o4 = reinterpret_cast<u32*>(out + i);
#pragma unroll 1
      for (u32 k = 0; k < steps; k += 2) {
        br.refill_point();  // at most 2 x 48 bits between points
        u32 hi, lo, used;
        br.window(hi, lo);
        *o4++ = huf_group4(table, lv, hi, lo, &used);
        br.p -= (i32)used;
        if (k + 1 < steps) {
          br.window(hi, lo);
          *o4++ = huf_group4(table, lv, hi, lo, &used);
          br.p -= (i32)used;
        }
      }
      i += 4 * steps;
      if (n - i < 4) {  // tail, then the stream must be exactly used up
        br.refill_point();
        for (; i < n; i++) {
          u32 hi, lo;
          br.window(hi, lo);
          const u32 e = huf_lookup(table, lv, hi);
          out[i] = (u8)e;
          br.p -= (i32)(e >> 8);
        }
        laneDone = true;
        if (br.p != br.b0) ctxs[frame].status = (u32)ZE_CORRUPTION;
      }
    }
    __syncwarp();
    u32 dm = __ballot_sync(kFull, laneDone);
    if (!quadIdle && ((dm >> (lane & ~3u)) & 0xFu) == 0xFu) quadIdle = true;
  }
}

// ------------------------------------------------------------------------------------------
// summary[0] = lowest failing frame index, summary[1] = frames that still have blocks to decode.
__global__ void k_frame_finish(const u8* __restrict__ src, const u8* __restrict__ dst, const FrameDesc* __restrict__ descs,
                               FrameCtx* __restrict__ ctxs, u32 nFrames, u32* __restrict__ summary) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  u32 i = t >> 2, q = t & 3;
  bool live = i < nFrames;
  FrameCtx* c = live ? &ctxs[i] : nullptr;
  u32 status = live ? c->status : 0;
  u32 flags = live ? c->flags : 0;
  bool hash = live && !status && (flags & FF_DONE) && !(flags & FF_FINISHED) && (flags & FF_CHECKSUM);
  u64 acc = xxh_init_acc(q);
  u32 len = 0;
  const u8* p = nullptr;
  if (hash) {
    len = c->dstPos;
    p = dst + descs[i].dstOff;
    u32 stripes = len >> 5;
    if (((uintptr_t)p & 7) == 0) {
      // 16 independent loads in flight before the dependent rounds: with few, large frames (256 KiB: 8192 rounds per
      // lane) this loop is load-latency bound, not bandwidth bound
      const u64* w = reinterpret_cast<const u64*>(p) + q;
      u32 k = 0;
      for (; k + 16 <= stripes; k += 16) {
        u64 v[16];
#pragma unroll
        for (u32 j = 0; j < 16; j++)  // volatile asm: the compiler otherwise sinks most of the loads between the rounds
          asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v[j]) : "l"(w + 4 * (u64)(k + j)));
#pragma unroll
        for (u32 j = 0; j < 16; j++) acc = xxh_round(acc, v[j]);
      }
      for (; k < stripes; k++) acc = xxh_round(acc, w[4 * (u64)k]);
    } else {
      const u8* b = p + 8 * q;
      for (u32 k = 0; k < stripes; k++) acc = xxh_round(acc, ld64(b + 32 * (u64)k));
    }
  }
  // gather the quad's accumulators in its lane 0 (all 32 lanes take part in the shuffles)
  u32 lane = threadIdx.x & 31, q0 = lane & ~3u;
  u64 v1 = __shfl_sync(kFull, acc, q0), v2 = __shfl_sync(kFull, acc, q0 + 1), v3 = __shfl_sync(kFull, acc, q0 + 2),
      v4 = __shfl_sync(kFull, acc, q0 + 3);
  if (!live || q != 0) return;
  if (!status && (flags & FF_DONE) && !(flags & FF_FINISHED)) {
    const FrameDesc d = descs[i];
    u32 tail = (flags & FF_CHECKSUM) ? 4u : 0u;
    if (c->srcPos + tail != d.srcLen) status = (flags & FF_CHECKSUM) && c->srcPos + tail > d.srcLen ? ZE_CHECKSUM_WRONG : ZE_SRC_WRONG;
    else if (d.exact && c->dstPos != d.dstCap) status = ZE_CORRUPTION;
    else if (c->fcs != ~0ull && c->fcs != c->dstPos) status = ZE_CORRUPTION;
    else if (hash) {
      u64 h;
      if (len >= 32) {
        h = xxh_rotl(v1, 1) + xxh_rotl(v2, 7) + xxh_rotl(v3, 12) + xxh_rotl(v4, 18);
        h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
      } else {
        h = kXP5;
      }
      h = xxh_finish(h, len, p + (len & ~31u), len & 31u);
      if ((u32)h != ld32(src + d.srcOff + c->srcPos)) status = ZE_CHECKSUM_WRONG;
    }
    c->flags = flags | FF_FINISHED;
    if (status) c->status = status;
  }
  if (status) atomicMin(&summary[0], i);
  else if (!(flags & FF_DONE)) atomicAdd(&summary[1], 1u);
}

// ------------------------------------------------------------------------------------------
static inline u32 div_up(u64 a, u32 b) { return (u32)((a + b - 1) / b); }

const char* kernel_name(int id) {
  static const char* names[K_COUNT] = {"start", "k_build_descs", "k_block_setup", "k_huf_decode", "k_seq_decode", "k_seq_execute",
                                       "k_frame_finish"};
  return id >= 0 && id < K_COUNT ? names[id] : "?";
}

KernelTimer::~KernelTimer() {
  for (size_t i = 0; i < poolCap_; i++) cudaEventDestroy(pool_[i]);
  free(pool_);
  free(marks_);
}
void KernelTimer::reset() {
  for (int i = 0; i < K_COUNT; i++) { ms[i] = 0; launches[i] = 0; }
  used_ = 0;
}
void KernelTimer::mark(int id, cudaStream_t st) {
  if (used_ == cap_) {
    size_t ncap = cap_ ? cap_ * 2 : 256;
    marks_ = static_cast<Mark*>(realloc(marks_, ncap * sizeof(Mark)));
    pool_ = static_cast<cudaEvent_t*>(realloc(pool_, ncap * sizeof(cudaEvent_t)));
    for (size_t i = poolCap_; i < ncap; i++) cudaEventCreate(&pool_[i]);
    poolCap_ = cap_ = ncap;
  }
  marks_[used_].id = id;
  marks_[used_].ev = pool_[used_];
  cudaEventRecord(pool_[used_], st);
  used_++;
}
void KernelTimer::collect() {
  for (size_t i = 1; i < used_; i++) {
    if (marks_[i].id == K_START) continue;
    float t = 0;
    if (cudaEventElapsedTime(&t, marks_[i - 1].ev, marks_[i].ev) == cudaSuccess) {
      ms[marks_[i].id] += t;
      launches[marks_[i].id]++;
    }
  }
  used_ = 0;
}
void KernelTimer::dump_timeline(FILE* f) {
  for (size_t i = 0; i < used_; i++) {
    float t = -1;
    cudaEventSynchronize(marks_[i].ev);
    cudaEventElapsedTime(&t, marks_[0].ev, marks_[i].ev);
    fprintf(f, "TL %zu %s %.4f\n", i, kernel_name(marks_[i].id), t);
  }
}
#define ZRA_MARK(id) do { if (timer) timer->mark(id, st); } while (0)

static int sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

static void configure_kernels() {
  // per device; cheap enough to repeat on every launch sequence
  cudaFuncSetAttribute(k_seq_decode<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SeqGeom<false>::kSmem);
  cudaFuncSetAttribute(k_seq_decode<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SeqGeom<true>::kSmem);
  cudaFuncSetAttribute(k_huf_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHufWarpSmem);
  cudaFuncSetAttribute(k_block_setup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSetupSmem);
}

size_t decode_scratch_bytes(u32 nFrames, u32 maxDstCap, DecodeLayout* lay) {
  u32 blk = maxDstCap < kBlockSizeMax ? maxDstCap : kBlockSizeMax;
  lay->litStride = (blk + 15u) & ~15u;
  if (lay->litStride == 0) lay->litStride = 16;
  lay->seqStride = blk / 3 + 1;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
  lay->offDescs = take(sizeof(FrameDesc) * (size_t)nFrames);
  lay->offCtxs = take(sizeof(FrameCtx) * (size_t)nFrames);
  lay->offTabs = take(sizeof(FrameTables) * (size_t)nFrames);
  lay->offLit = take((size_t)lay->litStride * nFrames);
  lay->offSeqs = take(sizeof(u64) * (size_t)lay->seqStride * nFrames);
  lay->offSummary = take(64);
  lay->offWork = take(sizeof(RoundWork));
  lay->offHufList = take(sizeof(u32) * (size_t)nFrames);
  lay->offSeqList = take(sizeof(u32) * (size_t)nFrames);
  lay->offRedoList = take(sizeof(u32) * (size_t)nFrames);
  return off;
}

void launch_build_descs(const void* archive, u64 tableOff, u64 headerSize, u64 archiveSize, u64 uncompressedSize, u32 frameSize,
                        u32 firstFrame, u32 nFrames, u64 dstBase, void* scratch, const DecodeLayout& lay, cudaStream_t st,
                        KernelTimer* timer) {
  if (!nFrames) return;
  u8* s = static_cast<u8*>(scratch);
  ZRA_MARK(K_START);
  k_build_descs<<<div_up(nFrames, 128), 128, 0, st>>>(static_cast<const u8*>(archive), tableOff, headerSize, archiveSize,
                                                      uncompressedSize, frameSize, firstFrame, nFrames, dstBase,
                                                      reinterpret_cast<FrameDesc*>(s + lay.offDescs),
                                                      reinterpret_cast<u32*>(s + lay.offSummary));
  ZRA_MARK(K_BUILD_DESCS);
}

void launch_summary_reset(void* scratch, const DecodeLayout& lay, cudaStream_t st) {
  // summary = {first failing frame, frames not finished, first bad seek-table entry, unused}
  cudaMemsetAsync(static_cast<u8*>(scratch) + lay.offSummary, 0xFF, 16, st);
  cudaMemsetAsync(static_cast<u8*>(scratch) + lay.offSummary + 4, 0, 4, st);
}

u32 launch_decode_rounds(const void* src, void* dst, u32 nFrames, u32 rounds, bool first, void* scratch, const DecodeLayout& lay,
                         cudaStream_t st, KernelTimer* timer, const SideLane* side) {
  if (!nFrames) return 0;
  configure_kernels();
  u8* s = static_cast<u8*>(scratch);
  auto* descs = reinterpret_cast<FrameDesc*>(s + lay.offDescs);
  auto* ctxs = reinterpret_cast<FrameCtx*>(s + lay.offCtxs);
  auto* tabs = reinterpret_cast<FrameTables*>(s + lay.offTabs);
  u8* lit = s + lay.offLit;
  u64* seqs = reinterpret_cast<u64*>(s + lay.offSeqs);
  auto* work = reinterpret_cast<RoundWork*>(s + lay.offWork);
  u32* hufList = reinterpret_cast<u32*>(s + lay.offHufList);
  u32* seqList = reinterpret_cast<u32*>(s + lay.offSeqList);
  const u8* in = static_cast<const u8*>(src);
  const u32 sms = (u32)sm_count();
  // persistent grids: as many warps as fit the SMs' shared memory, never more than there is work
  const u32 hufWarps = sms * 20 < div_up(nFrames, 8) ? sms * 20 : div_up(nFrames, 8);
  // sequence stage: one-warp CTAs of 32 table slots (84.5 / 52.5 KiB of shared memory), persistent, pulling frames from
  // the round's list; never more CTAs than the SMs hold at once, never more than there is work
  static const u32 seqPerSm = [] { const char* e = getenv("ZRA_B200_SEQ_CTAS_PER_SM"); return e ? (u32)atoi(e) : 0u; }();
  const u32 perSm = seqPerSm ? seqPerSm : 2u, perSmS = seqPerSm ? seqPerSm : 4u;
  const u32 seqCtas = sms * perSm < div_up(nFrames, 32) ? sms * perSm : div_up(nFrames, 32);
  const u32 seqCtasS = sms * perSmS < div_up(nFrames, 32) ? sms * perSmS : div_up(nFrames, 32);
  u32* redoList = reinterpret_cast<u32*>(s + lay.offRedoList);
  // frames of at most 32 KiB have fewer than 2048 sequences per block far more often than not: they get the small
  // geometry first and the general kernel only sweeps up what did not qualify (usually nothing: its CTAs exit at once)
  static const bool noSmall = getenv("ZRA_B200_NO_SMALL_SEQ") != nullptr;
  const u32 splitSmall = (lay.litStride <= (32u << 10) && !noSmall) ? 1u : 0u;
  for (u32 r = 0; r < rounds; r++) {
    cudaMemsetAsync(work, 0, sizeof(RoundWork), st);
    ZRA_MARK(K_START);
    k_block_setup<<<div_up(nFrames, 32), 32, kSetupSmem, st>>>(in, descs, ctxs, tabs, nFrames, (first && r == 0) ? 1u : 0u, work, hufList,
                                                               seqList, splitSmall);
    ZRA_MARK(K_BLOCK_SETUP);
    // the Huffman stage runs BESIDE the sequence stage on the side lane: both depend only on the block setup and both
    // are latency-bound (1 GiB of 64 KiB frames, 2 chunks: 5.46 -> 5.19 ms per step, gpurun_out/r03l)
    if (side) {
      cudaEventRecord(side->fork, st);
      cudaStreamWaitEvent(side->st, side->fork, 0);
      k_huf_decode<<<hufWarps, 32, kHufWarpSmem, side->st>>>(in, descs, ctxs, tabs, lit, lay.litStride, work, hufList);
      cudaEventRecord(side->join, side->st);
    } else {
      k_huf_decode<<<hufWarps, 32, kHufWarpSmem, st>>>(in, descs, ctxs, tabs, lit, lay.litStride, work, hufList);
    }
    ZRA_MARK(K_HUF_DECODE);
    if (splitSmall) k_seq_decode<true><<<seqCtasS, 32, SeqGeom<true>::kSmem, st>>>(in, descs, ctxs, tabs, seqs, lay.seqStride, work, seqList, redoList, nFrames);
    k_seq_decode<false><<<seqCtas, 32, SeqGeom<false>::kSmem, st>>>(in, descs, ctxs, tabs, seqs, lay.seqStride, work, seqList, redoList, nFrames);
    k_seq_redo<<<div_up(nFrames, 64), 64, 0, st>>>(in, descs, ctxs, tabs, seqs, lay.seqStride, work, redoList);
    ZRA_MARK(K_SEQ_DECODE);
    if (side) cudaStreamWaitEvent(st, side->join, 0);
    k_seq_execute<<<div_up((u64)nFrames * 32, kExecWarps * 32), kExecWarps * 32, 0, st>>>(in, static_cast<u8*>(dst), descs, ctxs, lit, lay.litStride, seqs,
                                                                 lay.seqStride, nFrames);
    ZRA_MARK(K_SEQ_EXECUTE);
  }
  return rounds * (splitSmall ? 6u : 5u);
}

void launch_frame_finish(const void* src, const void* dst, u32 nFrames, void* scratch, const DecodeLayout& lay, cudaStream_t st,
                         KernelTimer* timer) {
  if (!nFrames) return;
  u8* s = static_cast<u8*>(scratch);
  // "not finished" is recounted by every finish pass
  cudaMemsetAsync(s + lay.offSummary + 4, 0, 4, st);
  ZRA_MARK(K_START);
  k_frame_finish<<<div_up((u64)nFrames * 4, 128), 128, 0, st>>>(static_cast<const u8*>(src), static_cast<const u8*>(dst),
                                                                reinterpret_cast<const FrameDesc*>(s + lay.offDescs),
                                                                reinterpret_cast<FrameCtx*>(s + lay.offCtxs), nFrames,
                                                                reinterpret_cast<u32*>(s + lay.offSummary));
  ZRA_MARK(K_FRAME_FINISH);
}

u32 frame_status_offset() { return (u32)offsetof(FrameCtx, status); }
u32 frame_ctx_size() { return (u32)sizeof(FrameCtx); }
u32 frame_desc_size() { return (u32)sizeof(FrameDesc); }

}  // namespace zrab
