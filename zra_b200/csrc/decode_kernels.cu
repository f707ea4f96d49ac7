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
#include "decode_launch.h"
#include "xxh64.cuh"

namespace zrab {

constexpr u32 kFull = 0xFFFFFFFFu;
constexpr u32 kLongCopy = 32;  // copies at least this long are done by the whole warp
constexpr u32 kNone = 0xFFFFFFFFu;

// Per-round work lists, filled by k_block_setup.
struct RoundWork {
  u32 hufCount, seqCount;  // entries appended this round
  u32 hufNext, seqNext;    // consumer cursors
  u32 seqCountS, seqNextS; // the small-table frames (taken from the back of the sequence list)
};

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
constexpr u32 kSeqSmallLogMax = 8;     // LL and ML table logs up to this take the small sequence-stage geometry (below)
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
// FSE sequences. One lane per frame; the warp's 32 frames decode one sequence per iteration in
// lock-step with their three compact FSE tables in shared memory (2.5 KiB per frame, 88 frames per
// SM). The step itself (decode_core.cuh: seq_step) is branch-free; the inner loop runs a
// warp-uniform number of steps (the minimum left over the active lanes), so there is no per-step
// completion test. A lane that finishes its frame pulls the next one from the work list and the
// warp copies that frame's tables in cooperatively.
// Two geometries: the general one holds full-size tables (LL 512 + ML 512 + OF 256 entries = 2.5 KiB per frame:
// 88 frames per SM, three warps); the small one serves frames whose LL and ML table logs are at most 8 — every frame
// of fewer than 2048 sequences, i.e. 16 KiB frames (FSE_optimalTableLog, zstd/compress/fse_compress.c:325-342) — with
// 3 x 256 entries = 1.5 KiB per frame: 144 frames per SM, five warps. k_block_setup sorts the frames into the two
// work lists (the small list grows from the back of the same array).
template <bool SMALL, u32 SLOTS = 0>
struct SeqGeom {
  static constexpr u32 kLL = SMALL ? 256 : 512, kML = SMALL ? 256 : 512, kOF = 256;
  static constexpr u32 kEntries = kLL + kML + kOF;
  static constexpr u32 kSlots = SLOTS ? SLOTS : (SMALL ? 144 : 88);
  static constexpr u32 kThreads = (kSlots + 31) / 32 * 32;  // lanes >= kSlots idle
  static constexpr u32 kSmem = kSlots * kEntries * sizeof(CSym) + 128 * sizeof(u32) + kRingWords * kThreads * sizeof(u32);
};
static_assert(SeqGeom<false>::kSmem <= 227 * 1024 && SeqGeom<true>::kSmem <= 227 * 1024, "k_seq_decode shared memory");

template <bool SMALL, u32 SLOTS = 0>
__global__ void __launch_bounds__(SeqGeom<SMALL, SLOTS>::kThreads) k_seq_decode(const u8* __restrict__ src, const FrameDesc* __restrict__ descs,
                                                   FrameCtx* __restrict__ ctxs, const FrameTables* __restrict__ tabs,
                                                   u64* __restrict__ seqs, u32 seqStride, RoundWork* __restrict__ work,
                                                   const u32* __restrict__ seqList, u32 nFrames) {
  using G = SeqGeom<SMALL, SLOTS>;
  constexpr u32 kSeqSlots = G::kSlots, kSeqThreads = G::kThreads, kSeqSlotEntries = G::kEntries;
  extern __shared__ __align__(16) u8 smem[];
  CSym* slots = reinterpret_cast<CSym*>(smem);
  u32* lutLL = reinterpret_cast<u32*>(smem + kSeqSlots * kSeqSlotEntries * sizeof(CSym));
  u32* lutML = lutLL + 64;
  u32* ring = lutML + 64 + threadIdx.x;  // [kRingWords][kSeqThreads]: a warp access never conflicts
  const u32 lane = threadIdx.x & 31, slot = threadIdx.x, slotBase = threadIdx.x & ~31u;
  for (u32 k = threadIdx.x; k < 64; k += kSeqThreads) {
    lutLL[k] = k < 36 ? ll_lut(k) : 0u;
    lutML[k] = k < 53 ? ml_lut(k) : 0u;
  }
  __syncthreads();  // the only block-wide barrier: from here on the warps run independently
  const CSym* tLL = slots + (slot < kSeqSlots ? slot : 0) * kSeqSlotEntries;
  const CSym* tML = tLL + G::kLL;
  const CSym* tOF = tML + G::kML;
  const u32 total = SMALL ? work->seqCountS : work->seqCount;
  bool active = false, exhausted = slot >= kSeqSlots;
  u32 frame = kNone;
  SeqState st;
  st.i = 0; st.n = 0;
  u64* out = nullptr;
  for (;;) {
    // ---- idle lanes pull frames; the warp stages their tables
    if (__any_sync(kFull, !active && !exhausted)) {
      u32 f = kNone;
      if (!active && !exhausted) {
        u32 k = atomicAdd(SMALL ? &work->seqNextS : &work->seqNext, 1u);
        if (k < total) f = seqList[SMALL ? nFrames - 1 - k : k];
        else exhausted = true;
      }
      u32 got = __ballot_sync(kFull, f != kNone);
      while (got) {
        int who = __ffs(got) - 1;
        got &= got - 1;
        u32 wf = __shfl_sync(kFull, f, who);
        // FrameTables keeps full-size arrays (ll at 0, ml at 1 KiB, of at 2 KiB); a slot holds the first kLL / kML / kOF
        // entries of each, back to back (16 bytes = 8 entries per vector)
        const uint4* g = reinterpret_cast<const uint4*>(&tabs[wf]);
        uint4* d = reinterpret_cast<uint4*>(slots + (slotBase + who) * kSeqSlotEntries);
#pragma unroll
        for (u32 k = 0; k < kSeqSlotEntries / 8 / 32; k++) {
          const u32 v = lane + 32 * k;  // vector index inside the slot
          const u32 from = v < G::kLL / 8 ? v : (v < (G::kLL + G::kML) / 8 ? 64 + (v - G::kLL / 8) : 128 + (v - (G::kLL + G::kML) / 8));
          d[v] = __ldg(g + from);
        }
      }
      __syncwarp();
      if (f != kNone) {
        frame = f;
        out = seqs + (u64)f * seqStride;
        u32 err = seq_begin(src, descs[f], ctxs[f], seqStride, st, ring, kSeqThreads);
        if (err) {
          FrameCtx* g = &ctxs[f];
          if (!g->status) g->status = err;
          g->blkType = BT_NONE;
          g->flags |= FF_DONE;
        } else {
          active = true;
        }
      }
    }
    if (!__any_sync(kFull, active)) {
      if (__all_sync(kFull, exhausted)) break;
      continue;  // a lane whose frame failed to start fetches again
    }
    // ---- run until the first active lane reaches the end of its block
    const u32 steps = __reduce_min_sync(kFull, active ? st.n - st.i : 0xFFFFFFFFu);
    if (active) {
#pragma unroll 1
      for (u32 k = 0; k < steps; k += 2) {
        st.br.refill_point();  // at least every 2 steps (SeqReader)
        *out++ = seq_step(tLL, tML, tOF, lutLL, lutML, st);
        if (k + 1 < steps) *out++ = seq_step(tLL, tML, tOF, lutLL, lutML, st);
      }
    }
    __syncwarp();
    if (active && st.i == st.n) {
      FrameCtx* g = &ctxs[frame];
      FrameCtx tmp;
      u32 err = seq_end(st, tmp);
      if (!err) {
        g->rep[0] = tmp.rep[0]; g->rep[1] = tmp.rep[1]; g->rep[2] = tmp.rep[2];
        g->blkOut = tmp.blkOut; g->dstPos = tmp.dstPos;
      } else {
        if (!g->status) g->status = err;
        g->blkType = BT_NONE;
        g->flags |= FF_DONE;
      }
      active = false;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Whole-warp forward copy of n bytes between regions that do not overlap: 16-byte stores to the
// aligned body of dst, the source words funnel-shifted into place (src and dst may have any
// alignment). Reads whole aligned words, i.e. up to 3 bytes either side of the source range.
__device__ __forceinline__ void warp_copy_wide(u8* dst, const u8* src, u32 n, u32 lane) {
  u32 head = (16u - (u32)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
  if (head > n) head = n;
  if (lane < head) dst[lane] = src[lane];
  dst += head; src += head; n -= head;
  const u32 vecs = n >> 4;
  const u32 sh = (u32)(reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
  const u32* sw = reinterpret_cast<const u32*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
  uint4* dv = reinterpret_cast<uint4*>(dst);
  for (u32 v = lane; v < vecs; v += 32) {
    const u32* p = sw + 4 * v;
    u32 a = p[0], b = p[1], c = p[2], d = p[3], e = sh ? p[4] : 0u;
    uint4 o;
    o.x = __funnelshift_r(a, b, sh);
    o.y = __funnelshift_r(b, c, sh);
    o.z = __funnelshift_r(c, d, sh);
    o.w = __funnelshift_r(d, e, sh);
    dv[v] = o;
  }
  const u32 done = vecs << 4, tail = n & 15u;
  if (lane < tail) dst[done + lane] = src[done + lane];
}

// ---- lock-step short copies --------------------------------------------------------------
// Every lane copies its own n bytes (possibly 0); m is a warp-uniform upper bound of n. Four bytes
// per trip, written as predicated PTX (one predicate per byte position, loads before stores,
// immediate offsets): nvcc turns the equivalent C++ into nested divergent branches.
#define ZRA_PRED4 "setp.gt.s32 p0, %2, 0;\n\tsetp.gt.s32 p1, %2, 1;\n\tsetp.gt.s32 p2, %2, 2;\n\tsetp.gt.s32 p3, %2, 3;\n\t"
#define ZRA_COPY4(LD, ST)                                                                               \
  "{\n\t.reg .pred p0, p1, p2, p3;\n\t.reg .b32 b0, b1, b2, b3;\n\t" ZRA_PRED4                          \
  "@p0 " LD " b0, [%0];\n\t@p1 " LD " b1, [%0+1];\n\t@p2 " LD " b2, [%0+2];\n\t@p3 " LD " b3, [%0+3];\n\t" \
  "@p0 " ST " [%1], b0;\n\t@p1 " ST " [%1+1], b1;\n\t@p2 " ST " [%1+2], b2;\n\t@p3 " ST " [%1+3], b3;\n\t}"

// global (read-only data: literal scratch / input) -> global
__device__ __forceinline__ void lanes_copy_ro(u8* d, const u8* s, u32 n, u32 m) {
  i32 r = (i32)n;
#pragma unroll 1
  for (u32 k = 0; k < m; k += 4) {
    asm volatile(ZRA_COPY4("ld.global.nc.u8", "st.global.u8")::"l"(s), "l"(d), "r"(r) : "memory");
    s += 4; d += 4; r -= 4;
  }
}
// global (output written earlier by this warp) -> global
__device__ __forceinline__ void lanes_copy_gg(u8* d, const u8* s, u32 n, u32 m) {
  i32 r = (i32)n;
#pragma unroll 1
  for (u32 k = 0; k < m; k += 4) {
    asm volatile(ZRA_COPY4("ld.global.u8", "st.global.u8")::"l"(s), "l"(d), "r"(r) : "memory");
    s += 4; d += 4; r -= 4;
  }
}
// shared -> shared (32-bit shared-window addresses)
__device__ __forceinline__ void lanes_copy_ss(u32 d, u32 s, u32 n, u32 m) {
  i32 r = (i32)n;
#pragma unroll 1
  for (u32 k = 0; k < m; k += 4) {
    asm volatile(ZRA_COPY4("ld.shared.u8", "st.shared.u8")::"r"(s), "r"(d), "r"(r) : "memory");
    s += 4; d += 4; r -= 4;
  }
}
// generic (global or shared) -> shared, eight bytes per trip; the source is fetched as ALIGNED
// 32-bit words (one L1 request per 4 bytes instead of four) and funnel-shifted into place. Only
// words that hold at least one wanted byte are read.
__device__ __forceinline__ u32 ld_word_if(const u32* w, bool p) {
  u32 v = 0;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.u32 %0, [%1];\n\t}" : "+r"(v) : "l"(w), "r"((u32)p) : "memory");
  return v;
}
__device__ __forceinline__ void lanes_copy_xs(u32 d, const u8* s, u32 n, u32 m) {
  i32 r = (i32)n;
  const u32 mis = (u32)(reinterpret_cast<uintptr_t>(s) & 3u);
  const u32 sh = mis * 8u;
  const i32 thr1 = 4 - (i32)mis, thr2 = 8 - (i32)mis;  // word k+1 / k+2 is needed iff more than thr1 / thr2 bytes are left
  const u32* w = reinterpret_cast<const u32*>(reinterpret_cast<uintptr_t>(s) & ~(uintptr_t)3);
  u32 cur = ld_word_if(w, r > 0);
#pragma unroll 1
  for (u32 k = 0; k < m; k += 8) {
    const u32 w1 = ld_word_if(w + 1, r > thr1);
    const u32 w2 = ld_word_if(w + 2, r > thr2);
    const u32 x0 = __funnelshift_r(cur, w1, sh), x1 = __funnelshift_r(w1, w2, sh);
    asm volatile(
        "{\n\t.reg .pred p0, p1, p2, p3, p4, p5, p6, p7;\n\t.reg .b32 b1, b2, b3, b5, b6, b7;\n\t"
        "setp.gt.s32 p0, %3, 0;\n\tsetp.gt.s32 p1, %3, 1;\n\tsetp.gt.s32 p2, %3, 2;\n\tsetp.gt.s32 p3, %3, 3;\n\t"
        "setp.gt.s32 p4, %3, 4;\n\tsetp.gt.s32 p5, %3, 5;\n\tsetp.gt.s32 p6, %3, 6;\n\tsetp.gt.s32 p7, %3, 7;\n\t"
        "shr.u32 b1, %0, 8;\n\tshr.u32 b2, %0, 16;\n\tshr.u32 b3, %0, 24;\n\t"
        "shr.u32 b5, %1, 8;\n\tshr.u32 b6, %1, 16;\n\tshr.u32 b7, %1, 24;\n\t"
        "@p0 st.shared.u8 [%2], %0;\n\t@p1 st.shared.u8 [%2+1], b1;\n\t@p2 st.shared.u8 [%2+2], b2;\n\t@p3 st.shared.u8 [%2+3], b3;\n\t"
        "@p4 st.shared.u8 [%2+4], %1;\n\t@p5 st.shared.u8 [%2+5], b5;\n\t@p6 st.shared.u8 [%2+6], b6;\n\t@p7 st.shared.u8 [%2+7], b7;\n\t}"
        ::"r"(x0), "r"(x1), "r"(d), "r"(r) : "memory");
    cur = w2; w += 2; d += 8; r -= 8;
  }
}

__device__ __forceinline__ u64 shfl64(u64 v, int srcLane) {
  u32 lo = __shfl_sync(kFull, (u32)v, srcLane), hi = __shfl_sync(kFull, (u32)(v >> 32), srcLane);
  return (u64)lo | ((u64)hi << 32);
}
__device__ __forceinline__ u64 shfl64_up1(u64 v) {
  u32 lo = __shfl_up_sync(kFull, (u32)v, 1), hi = __shfl_up_sync(kFull, (u32)(v >> 32), 1);
  return (u64)lo | ((u64)hi << 32);
}

// Sequence execution, one warp per frame, 32 sequences per iteration (one per lane). The records
// are cumulative (decode_core.cuh), so a lane gets its literal source, output position and lengths
// from its own record and its left neighbour's. Literals never depend on matches; matches are
// resolved in rounds: everything below the first pending match is final, so that match can always
// run, and so can every later match whose source lies entirely below it.
//
// The L1 request rate, not the instruction count, bounds a byte-granular LZ copy (profiles/r01b:
// one request per byte moved), so a group's output is ASSEMBLED IN SHARED MEMORY: each warp owns a
// 4 KiB tile laid out at the same 16-byte phase as the destination; literals and far match sources
// are fetched as aligned 32-bit words, near matches copy tile to tile, and the finished group
// leaves with 16-byte coalesced stores. Groups that contain a long literal run or match (>= 64
// bytes) or regenerate more than the tile holds take the direct global path instead.
// Reference semantics: ZSTD_execSequence, zstd/decompress/zstd_decompress_block.c:704-793.
constexpr u32 kExecWarps = 8;
constexpr u32 kTileBytes = 4096;
constexpr u32 kTileStride = kTileBytes + 32;
constexpr u32 kShortMax = 64;  // sequences with ll and ml below this go through the tile

// Tried and measured slower (profiles/r01f): prefetching the next group's literals / match sources into L1
// (2.58 -> 2.69 .. 2.74 ms) and 5 CTAs per SM at 48 registers (3.07 ms).
__global__ void __launch_bounds__(kExecWarps * 32, 4) k_seq_execute(const u8* __restrict__ src, u8* dst, const FrameDesc* __restrict__ descs,
                                                     const FrameCtx* __restrict__ ctxs, const u8* __restrict__ lit, u32 litStride,
                                                     const u64* __restrict__ seqs, u32 seqStride, u32 nFrames) {
  __shared__ __align__(16) u8 tiles[kExecWarps][kTileStride];
  u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  u32 lane = threadIdx.x & 31;
  if (warp >= nFrames) return;
  const FrameCtx& c = ctxs[warp];
  if (c.status || c.blkType == BT_NONE) return;
  const FrameDesc d = descs[warp];
  u8* frame = dst + d.dstOff;  // frame-relative positions index this
  const u8* fsrc = src + d.srcOff;
  const u32 blkDst = c.blkDst;
  if (c.blkType == BT_RAW) {
    warp_copy_wide(frame + blkDst, fsrc + c.blkSrc, c.blkSize, lane);
    return;
  }
  if (c.blkType == BT_RLE) {
    u8 v = fsrc[c.blkSrc];
    for (u32 i = lane; i < c.blkSize; i += 32) frame[blkDst + i] = v;
    return;
  }
  // ---- compressed block
  u8* tile = tiles[threadIdx.x >> 5];
  const u32 tileS = (u32)__cvta_generic_to_shared(tile);
  const bool rle = c.litMode == LIT_RLE;
  const u8 rleByte = (u8)c.litSrc;
  const u8* litp = c.litMode == LIT_HUF ? lit + (u64)warp * litStride : fsrc + c.litSrc;
  const u64* sq = seqs + (u64)warp * seqStride;
  const u32 nbSeq = c.nbSeq;
  u8* blk = frame + blkDst;  // block-relative positions (the records' outEnd) index this
  u64 carry = 0;  // record of the last sequence of the previous iteration
  u32 pend = 0;   // bytes of the last tile group's partial final vector, kept in tile[0, pend) (warp-uniform)
  u8* pendG = nullptr;  // where they belong in the output buffer
  // lanes past the end repeat the last record (ll = ml = 0); the next group's records are requested a
  // whole iteration ahead
  u64 sNext = nbSeq ? __ldg(sq + (lane < nbSeq ? lane : nbSeq - 1)) : 0ull;
  for (u32 base = 0; base < nbSeq; base += 32) {
    const u64 s = sNext;
    {
      const u32 nidx = base + 32 + lane;
      if (base + 32 < nbSeq) sNext = __ldg(sq + (nidx < nbSeq ? nidx : nbSeq - 1));
    }
    u64 p = shfl64_up1(s);
    if (lane == 0) p = carry;
    const u32 S0 = rec_out_end(carry);  // block-relative start of this group's output
    carry = shfl64(s, 31);
    const u32 S = rec_out_end(carry) - S0;
    const u32 pl = rec_lit_end(p), po = rec_out_end(p);
    const u32 ll = rec_lit_end(s) - pl;
    const u32 ml = rec_out_end(s) - po - ll;
    const u32 off = rec_off(s);
    const bool viaTile = S <= kTileBytes && !__any_sync(kFull, ll >= kShortMax || ml >= kShortMax);
    if (viaTile) {
      // tile byte t <-> block byte S0 - a + t, a = 16-byte phase of the group's first output byte. The bytes of the
      // LAST, partial 16-byte vector of a group are not stored: they stay in the tile and become tile[0, a) of the
      // next group (`pend` of them, then a == pend), so that groups leave as whole 16-byte vectors only.
      const u32 a = (u32)(reinterpret_cast<uintptr_t>(blk + S0) & 15u);
      u8* gbase = blk + ((i32)S0 - (i32)a);  // 16-byte aligned; tile byte t is gbase[t]
      const bool headValid = pend != 0;      // tile[0, a) holds the carried bytes (not yet in the output buffer)
      const i32 lowT = headValid ? 0 : (i32)a;  // tile offsets from here on are in the tile, below it in the output buffer
      const u32 tl = po - S0 + a;  // tile offset of this lane's literals
      const u32 tm = tl + ll;      // ... and of its match
      // ---- literals
      {
        const u32 m = __reduce_max_sync(kFull, ll);
        if (rle) { for (u32 i = 0; i < ll; i++) tile[tl + i] = rleByte; }
        else lanes_copy_xs(tileS + tl, litp + pl, ll, m);
      }
      __syncwarp();
      // ---- matches: source offset relative to the tile
      const i32 ms = (i32)tm - (i32)off;
      bool pending = ml > 0;
      for (;;) {
        const u32 mask = __ballot_sync(kFull, pending);
        if (!mask) break;
        const int first = __ffs(mask) - 1;
        const i32 hwm = (i32)__shfl_sync(kFull, tm, first);
        const bool ready = pending && ((int)lane == first || ms + (i32)ml <= hwm);
        // one lock-step loop serves every ready lane whose source is entirely in the output buffer
        // or entirely in the tile (generic addresses); the rare rest is done after it
        const bool inTile = ms >= lowT, inOut = ms + (i32)ml <= lowT;
        const bool plain = ready && off >= ml && (inTile || inOut);
        const u8* sp = inTile ? tile + ms : gbase + ms;
        const u32 n = plain ? ml : 0;
        const u32 m = __reduce_max_sync(kFull, n);
        lanes_copy_xs(tileS + tm, sp, n, m);
        if (ready && !plain) {
          // straddles the tile start and / or overlaps its own output: byte-serial
          for (u32 i = 0; i < ml; i++) {
            const i32 t = ms + (i32)i;
            tile[tm + i] = t < lowT ? gbase[t] : tile[t];
          }
        }
        if (ready) pending = false;
        __syncwarp();
      }
      // ---- the group leaves: whole 16-byte vectors [firstFull, endA); the ragged start is written by bytes only when
      // no head was carried in (first tile group of a block, or after a direct group)
      {
        const u32 end = a + S, endA = end & ~15u;
        const u32 firstFull = headValid ? 0u : (a + 15u) & ~15u;
        if (!headValid && a) {
          const u32 i = a + lane, stop = firstFull < end ? firstFull : end;
          if (i < stop) gbase[i] = tile[i];
        }
        for (u32 lo = firstFull + 16u * lane; lo < endA; lo += 512u)
          *reinterpret_cast<uint4*>(gbase + lo) = *reinterpret_cast<const uint4*>(tile + lo);
        // the partial last vector is carried if this warp owns it from its first byte
        const u32 newPend = (end > endA && endA >= firstFull) ? end - endA : 0u;
        u8 keep = 0;
        if (lane < newPend) keep = tile[endA + lane];
        __syncwarp();
        if (lane < newPend) tile[lane] = keep;
        pend = newPend;
        pendG = gbase + endA;
      }
      __syncwarp();
      continue;
    }
    // ---- direct path (long runs / matches): straight to the output buffer, after the carried bytes
    if (pend) {
      if (lane < pend) pendG[lane] = tile[lane];
      pend = 0;
      __syncwarp();
    }
    const u32 myDst = blkDst + po;
    u32 longLit = __ballot_sync(kFull, ll >= kLongCopy);
    while (longLit) {
      int who = __ffs(longLit) - 1;
      longLit &= longLit - 1;
      u32 L = __shfl_sync(kFull, ll, who), from = __shfl_sync(kFull, pl, who), to = __shfl_sync(kFull, myDst, who);
      if (rle) { for (u32 i = lane; i < L; i += 32) frame[to + i] = rleByte; }
      else warp_copy_wide(frame + to, litp + from, L, lane);
    }
    {
      const u32 n = ll < kLongCopy ? ll : 0;
      const u32 m = __reduce_max_sync(kFull, n);
      if (rle) { for (u32 i = 0; i < n; i++) frame[myDst + i] = rleByte; }
      else lanes_copy_ro(frame + myDst, litp + pl, n, m);
    }
    __syncwarp();
    const u32 mpos = myDst + ll;
    const u32 msrc = mpos - off;
    bool pending = ml > 0;
    for (;;) {
      u32 mask = __ballot_sync(kFull, pending);
      if (!mask) break;
      int first = __ffs(mask) - 1;
      u32 hwm = __shfl_sync(kFull, mpos, first);
      u32 fml = __shfl_sync(kFull, ml, first);
      if (fml >= kLongCopy) {
        u32 fs = __shfl_sync(kFull, msrc, first), fo = __shfl_sync(kFull, off, first);
        if (fo >= fml) {
          warp_copy_wide(frame + hwm, frame + fs, fml, lane);
        } else if (fo >= 32) {
          // overlap further than a warp-width: 32-byte slices in order, each reads only final bytes
          for (u32 i = 0; i < fml; i += 32) {
            if (i + lane < fml) frame[hwm + i + lane] = frame[fs + i + lane];
            __syncwarp();
          }
        } else {
          // short period: the period [hwm-fo, hwm) is final, every byte is a lookup into it
          for (u32 i = lane; i < fml; i += 32) frame[hwm + i] = frame[fs + (i % fo)];
        }
        if ((int)lane == first) pending = false;
      } else {
        const bool ready = pending && ml < kLongCopy && ((int)lane == first || msrc + ml <= hwm);
        u32 n = ready ? ml : 0;
        if (ready && off < ml) {  // rare: short self-overlapping match (only the first pending one can be), byte-serial
          u32 j = 0;
          for (u32 i = 0; i < ml; i++) {
            frame[mpos + i] = frame[msrc + j];
            if (++j == off) j = 0;
          }
          n = 0;
        }
        const u32 m = __reduce_max_sync(kFull, n);
        lanes_copy_gg(frame + mpos, frame + msrc, n, m);
        if (ready) pending = false;
      }
      __syncwarp();
    }
  }
  if (pend) {
    if (lane < pend) pendG[lane] = tile[lane];
    __syncwarp();
  }
  // trailing literals
  const u32 litPos = rec_lit_end(carry);
  const u32 pos = blkDst + rec_out_end(carry);
  const u32 rest = c.litSize - litPos;
  if (rle) { for (u32 i = lane; i < rest; i += 32) frame[pos + i] = rleByte; }
  else warp_copy_wide(frame + pos, litp + litPos, rest, lane);
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
  cudaFuncSetAttribute(k_seq_decode<false, 72>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SeqGeom<false, 72>::kSmem);
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

void launch_decode_rounds(const void* src, void* dst, u32 nFrames, u32 rounds, bool first, void* scratch, const DecodeLayout& lay,
                          cudaStream_t st, KernelTimer* timer) {
  if (!nFrames) return;
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
  // Tried and measured slower (profiles/r01i): running the Huffman stage on a side stream beside the sequence stage
  // (with 78 / 72 / 64 slots to leave it shared memory): the Huffman warps take issue slots and shared-memory
  // bandwidth from the latency-critical sequence warps (6.43 -> 6.9 .. 7.8 ms per step).
  // Chunks of more than one wave (>= 148 x 88 frames: archives of 4 GiB and up) are throughput-bound, not latency-bound:
  // with 72 slots the sequence CTA leaves room for one execute CTA on its SM, and the two overlap (4 GiB: 174 -> 183 GB/s;
  // at 1 GiB the same choice costs 9 %: the execute warps slow the latency-critical chains down; profiles/r01v).
  static const u32 slotsEnv = [] { const char* e = getenv("ZRA_B200_SEQ_SLOTS"); return e ? (u32)atoi(e) : 0u; }();
  const u32 genSlots = slotsEnv == 72 || slotsEnv == 88 ? slotsEnv : (nFrames >= sms * SeqGeom<false>::kSlots ? 72u : SeqGeom<false>::kSlots);
  const u32 seqCtas = sms < div_up(nFrames, genSlots) ? sms : div_up(nFrames, genSlots);
  const u32 seqCtasS = sms < div_up(nFrames, SeqGeom<true>::kSlots) ? sms : div_up(nFrames, SeqGeom<true>::kSlots);
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
    k_huf_decode<<<hufWarps, 32, kHufWarpSmem, st>>>(in, descs, ctxs, tabs, lit, lay.litStride, work, hufList);
    ZRA_MARK(K_HUF_DECODE);
    if (splitSmall)
      k_seq_decode<true><<<seqCtasS, SeqGeom<true>::kThreads, SeqGeom<true>::kSmem, st>>>(in, descs, ctxs, tabs, seqs, lay.seqStride, work, seqList, nFrames);
    if (genSlots == 72) k_seq_decode<false, 72><<<seqCtas, SeqGeom<false, 72>::kThreads, SeqGeom<false, 72>::kSmem, st>>>(in, descs, ctxs, tabs, seqs, lay.seqStride, work, seqList, nFrames);
    else
    k_seq_decode<false><<<seqCtas, SeqGeom<false>::kThreads, SeqGeom<false>::kSmem, st>>>(in, descs, ctxs, tabs, seqs, lay.seqStride, work, seqList, nFrames);
    ZRA_MARK(K_SEQ_DECODE);
    k_seq_execute<<<div_up((u64)nFrames * 32, kExecWarps * 32), kExecWarps * 32, 0, st>>>(in, static_cast<u8*>(dst), descs, ctxs, lit, lay.litStride, seqs,
                                                                 lay.seqStride, nFrames);
    ZRA_MARK(K_SEQ_EXECUTE);
  }
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
