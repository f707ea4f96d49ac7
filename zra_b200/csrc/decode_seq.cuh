// decode_seq.cuh — k_seq_decode: FSE sequence decode, one lane per frame, 32 frames per warp in lock-step.
//
// Reference semantics: ZSTD_decodeSequence + the loop of ZSTD_decompressSequences_body,
// zstd/decompress/zstd_decompress_block.c:838-948, :952-1080 (bit order of a sequence: offset extra bits, match
// extra bits, literal extra bits, then the LL / ML / OF state updates; the last sequence updates no state).
//
// Every stage of this kernel is a serial chain per frame, and an archive offers only as many chains as it has
// frames (16 384 for 1 GiB of 64 KiB frames = 3.5 warps per SM): the kernel is bound by the LATENCY of one step, and
// a single warp per scheduler issues its dependent instructions about one every other cycle. So the step is built
// for few instructions and a short chain (profiles/r02i measured the previous step at ~190 SASS instructions, 17.5
// of 32 lanes active; this one is branch-free and about half as long):
//  * tables stay in the compact 2-byte form (symbol | nextStateCounter << 6, entropy.cuh) so that 32 frames fit in
//    80 KiB of shared memory; the new state is one funnel shift (counter:bits << nbBits) and a mask;
//  * the compressed bytes are staged with cp.async (LDGSTS) into a per-lane ring of 32 words laid out [word][lane]
//    — no register staging, no landing step, no scoreboard coupling between the 32 unrelated streams; a 96-bit
//    window is rebuilt from four ring words at the cursor every step (no refill branch), which covers the longest
//    legal sequence (31 + 16 + 16 extra bits + 26 state bits);
//  * the checks run as a sticky flag; a frame that trips it is decoded again by the careful thread-serial
//    seq_decode (decode_core.cuh) to find the exact status code the reference would report;
//  * the last sequence of a block (no state update) is peeled out of the lock-step loop.
// A CTA is ONE warp with its 32 table slots (so that the CTAs of several chunks, and the execute kernel's CTAs, share
// an SM), pulling frames from the round's work list as lanes finish.
#pragma once
#include "decode_core.cuh"

namespace zrab {

constexpr u32 kSeqFull = 0xFFFFFFFFu;
constexpr u32 kSeqNone = 0xFFFFFFFFu;
constexpr u32 kSeqSmallLogMax = 8;  // LL and ML table logs up to this take the small geometry
constexpr u32 kSeqRingWords = 32;   // per lane
constexpr u32 kSeqWarpLanes = 32;

// Per-round work lists, filled by k_block_setup.
struct RoundWork {
  u32 hufCount, seqCount;  // entries appended this round
  u32 hufNext, seqNext;    // consumer cursors
  u32 seqCountS, seqNextS; // the small-table frames (taken from the back of the sequence list)
  u32 redoCount, pad;      // frames the fast sequence loop flagged (decoded again by the careful path)
};

template <bool SMALL>
struct SeqGeom {
  static constexpr u32 kLL = SMALL ? 256 : 512, kML = SMALL ? 256 : 512, kOF = 256;
  static constexpr u32 kEntries = kLL + kML + kOF;
  // shared memory of one warp: [code LUTs, 256-byte aligned][3 mirror rows][ring: 32 rows of 128 bytes, 4 KiB aligned]
  // [32 table slots]. The alignments let the hot loop merge an index into an address with one LOP3.
  static constexpr u32 kRingBytes = kSeqRingWords * kSeqWarpLanes * 4;  // 4 KiB
  static constexpr u32 kTabBytes = kSeqWarpLanes * kEntries * (u32)sizeof(CSym);
  static constexpr u32 kSmem = 1280 + 4096 + kRingBytes + kTabBytes;    // LUTs + mirror rows (+ their alignment), ring alignment slack, ring, tables
};

// ---- shared-memory access by explicit address. On the GPU an SAddr is a 32-bit shared-window address and the loads
// are ld.shared with immediate offsets (no generic-to-shared conversion in the loop); in the host logic build it is a
// pointer.
#if defined(__CUDA_ARCH__)
typedef u32 SAddr;
ZRA_DEV SAddr s_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
template <int OFF>
ZRA_DEV u32 s_ld32(SAddr a) { u32 v; asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF) : "memory"); return v; }
ZRA_DEV u32 s_ld16(SAddr a) { u32 v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
#else
typedef uintptr_t SAddr;
ZRA_DEV SAddr s_addr(const void* p) { return reinterpret_cast<uintptr_t>(p); }
template <int OFF>
ZRA_DEV u32 s_ld32(SAddr a) { return *reinterpret_cast<const u32*>(a + OFF); }
ZRA_DEV u32 s_ld16(SAddr a) { return *reinterpret_cast<const u16*>(a); }
#endif
ZRA_DEV SAddr s_merge(SAddr base, u32 bits) { return base | bits; }  // base has zeros where `bits` may have ones

// Shared-memory addresses a lane keeps in registers.
struct SeqSm {
  SAddr ringTop;       // this lane's word of ring row 0 (4 KiB aligned + 4 * lane); row q at +128 q; rows -3..-1 mirror 29..31
  SAddr lutLL, lutML;  // packed (baseline | extra bits << 24) per code; 256-byte aligned
  SAddr tLL, tML, tOF; // this lane's three tables
};

template <int N>
ZRA_DEV void cp_async_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// Per-frame state of the fast loop (registers of the lane that owns the frame).
struct FastSeq {
  const u32* g0;          // 16-byte aligned address of word 0 of the stream's "word space"
  i32 p;                  // index of the next unread bit, plus one
  i32 b0;                 // index of the stream's first bit
  i32 reqW;               // lowest word index requested so far (multiple of 4)
  SAddr aLL, aML, aOF;    // current cells
  u32 rep0, rep1, rep2;
  u32 litUsed, produced;
  u32 i, n;
  u32 llLog, mlLog, ofLog;
  u32 litSize, room, blkDst;
  u32 bad;                // sticky: sign bit set once any check failed
};

// Requests (cp.async: LDGSTS, no register staging) the 4-word group below reqW if `on` and the cursor word k is within
// `lead` words of it (the caller commits). Never goes below word 0: bits under the stream's first group read as
// whatever the ring holds (an over-read is caught by p != b0 at the end). Word w lives in ring row w & 31; rows 29..31
// are written twice (mirror rows -3..-1), so that the four words below any cursor sit at fixed offsets from one address.
ZRA_DEV void ring_request(FastSeq& s, const SeqSm& sm, i32 k, i32 lead, bool on) {
  const bool go = on && s.reqW > 0 && k - s.reqW < lead;
  const i32 w = s.reqW - 4;
  const u32 row = (u32)w & (kSeqRingWords - 1);
#if defined(__CUDA_ARCH__)
  const SAddr dst = s_merge(sm.ringTop, row << 7);
  const u32* src = s.g0 + w;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %2, 0;\n\tsetp.eq.and.u32 q, %3, 28, p;\n\t"
      "@p cp.async.ca.shared.global [%0], [%1], 4;\n\t@p cp.async.ca.shared.global [%0+128], [%1+4], 4;\n\t"
      "@p cp.async.ca.shared.global [%0+256], [%1+8], 4;\n\t@p cp.async.ca.shared.global [%0+384], [%1+12], 4;\n\t"
      "@q cp.async.ca.shared.global [%0+-3968], [%1+4], 4;\n\t@q cp.async.ca.shared.global [%0+-3840], [%1+8], 4;\n\t"
      "@q cp.async.ca.shared.global [%0+-3712], [%1+12], 4;\n\t}"
      ::"r"(dst), "l"(src), "r"((u32)go), "r"(row) : "memory");
#else
  if (go) {
    for (u32 j = 0; j < 4; j++) {
      *reinterpret_cast<u32*>(sm.ringTop + 128 * (row + j)) = s.g0[w + (i32)j];
      if (row + j >= 29) *reinterpret_cast<u32*>(sm.ringTop + 128 * (row + j) - 4096) = s.g0[w + (i32)j];
    }
  }
#endif
  s.reqW = go ? w : s.reqW;
}

ZRA_DEV void cp_async_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}

// Refill point of the lock-step loops: at most kSeqPointSteps steps (4 x 89 bits: the cursor crosses at most 12 word
// boundaries, and the window reaches 3 words below it) are consumed between two points, so a point must leave every
// word from the cursor down to cursor - 15 landed. cp.async groups are tracked per WARP, and with 32 unrelated streams
// some lane asks for a group at almost every point: a lane keeps its requests 24..28 words below its cursor — one
// group per point covers the usual consumption of 1.5 .. 3 words — and the warp waits for all but the newest TWO
// groups: what a lane requested two points ago has had ~2600 cycles to land (the slowest of 32 DRAM requests is what the
// warp waits for: with one point of slack the wait was 60 cycles per step, profiles/r05), i.e. everything above
// reqW + 8 <= cursor - 16. A stream of long sequences (> 5 words a point) outruns one request per point; then (any lane
// closer than kSeqFloor words after its request) every such lane tops up and the warp waits for everything. (Two steps
// per point and "all but the newest three" before: the predicated-off request block was 18 % of the kernel's
// instructions, profiles/r03w.)
constexpr u32 kSeqPointSteps = 4;
#ifndef ZRA_SEQ_PENDING
#define ZRA_SEQ_PENDING 2
#endif
constexpr int kSeqPending = ZRA_SEQ_PENDING;               // groups that may still be in flight after a point
constexpr i32 kSeqLead = 24;                               // a request goes out while the cursor is closer than this to reqW (32-word ring: < 28 keeps every row above the cursor's window intact)
constexpr i32 kSeqFloor = 15 + 4 * kSeqPending;             // ... which leaves reqW + 4 * kSeqPending landed: the cursor must be this far above reqW
ZRA_DEV void refill_point(FastSeq& s, const SeqSm& sm, bool active) {
  const i32 kw = (s.p - 1) >> 5;
  ring_request(s, sm, kw, kSeqLead, active);
  if (__any_sync(kSeqFull, active && s.reqW > 0 && kw - s.reqW < kSeqFloor)) {
    for (u32 g = 0; g < 3; g++) ring_request(s, sm, kw, kSeqFloor, active);
    cp_async_commit();
    cp_async_wait<0>();
  } else {
    cp_async_commit();
    cp_async_wait<kSeqPending>();
  }
}

// The 96-bit window at the cursor: bit 31 of H is the next unread bit.
ZRA_DEV void seq_window(const SeqSm& sm, i32 p, u32& H, u32& M, u32& L) {
  const i32 t = p - 1;
  const u32 sh = ~(u32)t & 31u;
  const SAddr row = s_merge(sm.ringTop, ((u32)t << 2) & 0xF80u);  // row (t >> 5) & 31
  const u32 w0 = s_ld32<0>(row), w1 = s_ld32<-128>(row), w2 = s_ld32<-256>(row), w3 = s_ld32<-384>(row);
  H = fsh_lc(w1, w0, sh); M = fsh_lc(w2, w1, sh); L = fsh_lc(w3, w2, sh);
}

// Starts the fast loop on the current block of one frame. Returns false when the block cannot start (the caller then
// leaves the frame to the careful path).
ZRA_DEV bool fast_begin(const u8* srcBase, const FrameDesc& d, const FrameCtx& c, u32 seqCap, FastSeq& s, const SeqSm& sm) {
  const u32 len = c.seqLen;
  if (c.nbSeq > seqCap || len == 0 || d.dstCap >= (1u << 30)) return false;  // (the sticky checks are sign tests)
  const u64 byteOff = d.srcOff + c.seqOff;
  const u32 last = srcBase[byteOff + len - 1];
  if (last == 0) return false;
  const u8* first = srcBase + byteOff;
  const u8* grp = reinterpret_cast<const u8*>(reinterpret_cast<uintptr_t>(first) & ~(uintptr_t)15);
  s.g0 = reinterpret_cast<const u32*>(grp);
  s.b0 = (i32)(first - grp) * 8;
  s.p = s.b0 + (i32)((len - 1) * 8 + highbit32(last));
  // initial fill: the top seven groups (28 words), synchronously
  const i32 k = (s.p - 1 >= 0 ? s.p - 1 : 0) >> 5;
  s.reqW = (k & ~3) + 4;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (u32 g = 0; g < 7; g++) ring_request(s, sm, k, 1 << 30, true);
  cp_async_commit();
  cp_async_wait<0>();
  s.llLog = c.llLog; s.ofLog = c.ofLog; s.mlLog = c.mlLog;
  // initial states: LL, OF, ML in that order (zstd_decompress_block.c:1000-1003)
  {
    u32 hi, lo, lo2;
    seq_window(sm, s.p, hi, lo, lo2);
    const u32 a = win_take(hi, lo, s.llLog);
    const u32 b = win_take(hi, lo, s.ofLog);
    const u32 e = win_take(hi, lo, s.mlLog);
    s.aLL = sm.tLL + 2 * a;
    s.aOF = sm.tOF + 2 * b;
    s.aML = sm.tML + 2 * e;
    s.p -= (i32)(s.llLog + s.ofLog + s.mlLog);
  }
  s.rep0 = c.rep[0]; s.rep1 = c.rep[1]; s.rep2 = c.rep[2];
  s.litUsed = 0; s.produced = 0; s.i = 0; s.n = c.nbSeq;
  s.litSize = c.litSize;
  const u32 room = d.dstCap - c.blkDst;
  s.room = room < kBlockSizeMax ? room : kBlockSizeMax;
  s.blkDst = c.blkDst;
  s.bad = 0;
  return true;
}

// One sequence. LAST: the block's final sequence (no state update, its bits are not consumed).
template <bool LAST>
ZRA_DEV u64 fast_step(const SeqSm& sm, FastSeq& s) {
  u32 H, M, L;
  seq_window(sm, s.p, H, M, L);
  // ---- cells
  const u32 cLL = s_ld16(s.aLL), cML = s_ld16(s.aML), cOF = s_ld16(s.aOF);
  const u32 lutl = s_ld32<0>(s_merge(sm.lutLL, (cLL << 2) & 0xFCu)), lutm = s_ld32<0>(s_merge(sm.lutML, (cML << 2) & 0xFCu));
  const u32 ofc = cOF & 63u;
  const u32 llBits = lutl >> 24, mlBits = lutm >> 24;
  // ---- extra bits, in stream order: offset | match length | literal length
  const u32 ofx = fsh_lc(H, 0u, ofc);
  const u32 X = fsh_lc(M, H, ofc), X2 = fsh_lc(L, M, ofc), X3 = fsh_lc(0u, L, ofc);
  const u32 mlx = fsh_lc(X, 0u, mlBits);
  const u32 Y = fsh_lc(X2, X, mlBits), Y2 = fsh_lc(X3, X2, mlBits);
  const u32 llx = fsh_lc(Y, 0u, llBits);
  u32 used = ofc + mlBits + llBits;
  if (!LAST) {
    // ---- state updates: LL, ML, OF. new state = (counter : stream bits) << nbBits, minus the table size
    const u32 Z = fsh_lc(Y2, Y, llBits);
    const u32 nsLL = cLL >> 6, nsML = cML >> 6, nsOF = cOF >> 6;
    const u32 nbLL = s.llLog - highbit32(nsLL), nbML = s.mlLog - highbit32(nsML), nbOF = s.ofLog - highbit32(nsOF);
    const u32 Z2 = Z << nbLL;      // nbLL <= 9
    const u32 Z3 = Z2 << nbML;
    s.aLL = sm.tLL + 2 * (fsh_lc(Z, nsLL, nbLL) & ((1u << s.llLog) - 1u));
    s.aML = sm.tML + 2 * (fsh_lc(Z2, nsML, nbML) & ((1u << s.mlLog) - 1u));
    s.aOF = sm.tOF + 2 * (fsh_lc(Z3, nsOF, nbOF) & ((1u << s.ofLog) - 1u));
    used += nbLL + nbML + nbOF;
  }
  s.p -= (i32)used;
  // ---- lengths and offset (zstd_decompress_block.c:871-917)
  const u32 llBase = lutl & 0xFFFFFFu;
  const u32 ll = llBase + llx, ml = (lutm & 0xFFFFFFu) + mlx;
  // idx: 0..2 = repeat offsets (shifted by one when the literal length is 0), 3 = rep0 - 1, 4 = a new offset
  u32 idx = ofc + ofx + (llBase == 0 ? 1u : 0u);
  idx = ofc > 1u ? 4u : idx;
  const u32 newOff = (1u << (ofc & 31u)) - 3u + ofx;
  u32 r3 = s.rep0 - 1u;
  r3 += !r3;
  u32 offset = s.rep0;
  offset = idx >= 1u ? s.rep1 : offset;
  offset = idx >= 2u ? s.rep2 : offset;
  offset = idx >= 3u ? r3 : offset;
  offset = idx >= 4u ? newOff : offset;
  s.rep2 = idx >= 2u ? s.rep1 : s.rep2;
  s.rep1 = idx >= 1u ? s.rep0 : s.rep1;
  s.rep0 = offset;
  // ---- checks (sticky): offset reaches no further back than the frame's output so far; the block's output and its
  // literals stay inside their buffers. Wrapped values only ever show up after the flag is set.
  const u32 litEnd = s.litUsed + ll;
  const u32 outEnd = s.produced + ll + ml;
  s.bad |= (s.blkDst + s.produced + ll - offset) | (0u - (offset >> 28)) | (s.room - outEnd) | (s.litSize - litEnd);
  s.litUsed = litEnd;
  s.produced = outEnd;
  s.i++;
  return (u64)(litEnd & 0x3FFFFu) | ((u64)(outEnd & 0x3FFFFu) << 18) | ((u64)(offset & kMaxOffset) << 36);
}

// ------------------------------------------------------------------------------------------
// The kernel: one warp per CTA. Dynamic shared memory: SeqGeom<SMALL>::kSmem.
template <bool SMALL>
__global__ void __launch_bounds__(32) k_seq_decode(const u8* __restrict__ src, const FrameDesc* __restrict__ descs, FrameCtx* __restrict__ ctxs,
                                                   const FrameTables* __restrict__ tabs, u64* __restrict__ seqs, u32 seqStride,
                                                   RoundWork* __restrict__ work, const u32* __restrict__ seqList, u32* __restrict__ redoList,
                                                   u32 nFrames) {
  using G = SeqGeom<SMALL>;
  ZRA_DYN_SMEM(smemRaw);
  const u32 lane = threadIdx.x;
  // carve: LUTs at a 256-byte boundary, then the ring at the next 4 KiB boundary that leaves 384 bytes for the mirror rows
  u8* lutP = smemRaw + ((256u - (u32)(s_addr(smemRaw) & 255u)) & 255u);
  u8* ringP = lutP + 512 + 384;
  ringP += (4096u - (u32)(s_addr(ringP) & 4095u)) & 4095u;
  CSym* tab = reinterpret_cast<CSym*>(ringP + G::kRingBytes);
  u32* lutLL = reinterpret_cast<u32*>(lutP);
  u32* lutML = lutLL + 64;
  for (u32 k = lane; k < 64; k += 32) {
    lutLL[k] = k < 36 ? ll_lut(k) : 0u;
    lutML[k] = k < 53 ? ml_lut(k) : 0u;
  }
  __syncwarp();
  SeqSm sm;
  sm.ringTop = s_addr(ringP) + 4 * lane;
  sm.lutLL = s_addr(lutLL);
  sm.lutML = s_addr(lutML);
  sm.tLL = s_addr(tab + lane * G::kEntries);
  sm.tML = sm.tLL + 2 * G::kLL;
  sm.tOF = sm.tML + 2 * G::kML;
  const u32 total = SMALL ? work->seqCountS : work->seqCount;
  bool active = false, exhausted = false;
  u32 frame = kSeqNone;
  FastSeq st;
  st.i = 0; st.n = 0; st.p = 0; st.b0 = 0; st.reqW = 0; st.g0 = nullptr;
  st.aLL = sm.tLL; st.aML = sm.tML; st.aOF = sm.tOF;  // a lane without a frame steps too: its cells must be addresses
  st.llLog = st.mlLog = st.ofLog = 0;
  st.rep0 = st.rep1 = st.rep2 = 1; st.litUsed = st.produced = st.litSize = st.room = st.blkDst = st.bad = 0;
  u64* out = nullptr;
  for (;;) {
    // ---- idle lanes pull frames; the warp stages their tables
    if (__any_sync(kSeqFull, !active && !exhausted)) {
      u32 f = kSeqNone;
      if (!active && !exhausted) {
        u32 k = atomicAdd(SMALL ? &work->seqNextS : &work->seqNext, 1u);
        if (k < total) f = seqList[SMALL ? nFrames - 1 - k : k];
        else exhausted = true;
      }
      u32 got = __ballot_sync(kSeqFull, f != kSeqNone);
      while (got) {
        const int who = __ffs((int)got) - 1;
        got &= got - 1;
        const u32 wf = __shfl_sync(kSeqFull, f, who);
        // FrameTables keeps full-size arrays (ll at 0, ml at 1 KiB, of at 2 KiB); a slot holds the first kLL / kML / kOF
        // entries of each, back to back (16 bytes = 8 entries per vector)
        const uint4* g = reinterpret_cast<const uint4*>(&tabs[wf]);
        uint4* d = reinterpret_cast<uint4*>(tab + (u32)who * G::kEntries);
        for (u32 v = lane; v < G::kEntries / 8; v += 32) {
          const u32 from = v < G::kLL / 8 ? v : (v < (G::kLL + G::kML) / 8 ? 64 + (v - G::kLL / 8) : 128 + (v - (G::kLL + G::kML) / 8));
          d[v] = g[from];
        }
      }
      __syncwarp();
      if (f != kSeqNone) {
        frame = f;
        out = seqs + (u64)f * seqStride;
        if (fast_begin(src, descs[f], ctxs[f], seqStride, st, sm)) active = true;
        else redoList[atomicAdd(&work->redoCount, 1u)] = f;
      }
    }
    if (!__any_sync(kSeqFull, active)) {
      if (__all_sync(kSeqFull, exhausted)) break;
      continue;  // a lane whose frame could not start fetches again
    }
    // ---- lock-step until the first active lane is one sequence from the end of its block
    const u32 steps = __reduce_min_sync(kSeqFull, active ? st.n - st.i - 1u : 0xFFFFFFFFu);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (u32 k = 0; k < steps; k += kSeqPointSteps) {
      refill_point(st, sm, active);
      // every lane steps (a lane without a frame walks harmlessly through its own slot and ring): no branch, only the
      // record store is predicated
      const u64 r0 = fast_step<false>(sm, st);
      if (active) out[0] = r0;
      if (k + 1 < steps) {
        const u64 r1 = fast_step<false>(sm, st);
        if (active) out[1] = r1;
      }
      if (k + 2 < steps) {
        const u64 r2 = fast_step<false>(sm, st);
        if (active) out[2] = r2;
      }
      if (k + 3 < steps) {
        const u64 r3 = fast_step<false>(sm, st);
        if (active) out[3] = r3;
      }
      out += kSeqPointSteps;
    }
    if (active) out -= (kSeqPointSteps - (steps & (kSeqPointSteps - 1u))) & (kSeqPointSteps - 1u);
    // ---- lanes at their last sequence: no state update, then the end-of-block checks
    if (active && st.n - st.i == 1u) {
      ring_request(st, sm, (st.p - 1) >> 5, 24, true);
      cp_async_commit();
      cp_async_wait<0>();
      *out = fast_step<true>(sm, st);
      const u32 lastLits = st.litSize - st.litUsed;
      const bool ok = !(st.bad >> 31) && st.p == st.b0 && lastLits <= st.room - st.produced;
      if (ok) {
        FrameCtx* g = &ctxs[frame];
        g->rep[0] = st.rep0; g->rep[1] = st.rep1; g->rep[2] = st.rep2;
        g->blkOut = st.produced + lastLits;
        g->dstPos = st.blkDst + st.produced + lastLits;
      } else {
        redoList[atomicAdd(&work->redoCount, 1u)] = frame;  // the careful path finds the status code
      }
      active = false;
    }
  }
}

// Frames the fast loop flagged: decoded again by the careful thread-serial path (exact reference status codes).
__global__ void k_seq_redo(const u8* __restrict__ src, const FrameDesc* __restrict__ descs, FrameCtx* __restrict__ ctxs,
                           const FrameTables* __restrict__ tabs, u64* __restrict__ seqs, u32 seqStride, const RoundWork* __restrict__ work,
                           const u32* __restrict__ redoList) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= work->redoCount) return;
  const u32 f = redoList[i];
  seq_decode(src, descs[f], ctxs[f], tabs[f], seqs + (u64)f * seqStride, seqStride);
}

}  // namespace zrab
