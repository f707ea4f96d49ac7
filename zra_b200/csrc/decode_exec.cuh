// decode_exec.cuh — k_seq_execute: sequence execution (literal + match copies), one warp per frame.
//
// Reference semantics: ZSTD_execSequence, zstd/decompress/zstd_decompress_block.c:704-793 (literal copy, then an
// overlap-safe match copy from `offset` bytes back), driven by the record stream k_seq_decode leaves in HBM.
//
// The warp takes 32 sequences at a time (one cumulative record per lane, decode_core.cuh). A group's output is
// ASSEMBLED BY BYTE GATHER in a shared-memory tile that has the 16-byte phase of its destination:
//   1. every lane describes its two runs (literals, match) in a small run table: the 64-bit distance from a
//      destination byte to its source byte is constant inside a run; and marks the tile byte where each run starts;
//   2. the tile is walked in rows of 32 bytes, one byte per lane: the marks of a row (one ballot) give every byte its
//      run (popc), the run gives the source address; bytes whose source is final — literals, and matches that reach
//      below the tile, by far the most — are fetched with ONE coalesced byte load per row (32 neighbouring
//      destination bytes come from a handful of source runs, i.e. a handful of sectors) and stored into the tile;
//   3. bytes whose source lies inside the tile wait for pass two, which walks the rows that have such bytes in
//      order (a source always precedes its destination), tile to tile; bytes that depend on their own row resolve
//      in ballot rounds;
//   4. the group leaves as whole 16-byte vectors; the partial last vector is also written by bytes (so that the
//      output buffer is always complete below the cursor) and carried into the next group's tile.
// The work per byte does not depend on the sequence lengths, there is no per-sequence loop and nothing is
// serialised on the longest sequence of a group (the lane-per-sequence copy loops this replaces ran at 1.9
// warp-instructions per output byte, profiles/r02i). Groups with a long literal run or match (>= 64 bytes), with
// RLE literals, or larger than the tile take the direct path: long copies by the whole warp with 16-byte stores.
#pragma once
#include "decode_core.cuh"

namespace zrab {

constexpr u32 kExecFull = 0xFFFFFFFFu;
constexpr u32 kLongCopy = 32;    // direct path: copies at least this long are done by the whole warp
constexpr u32 kExecWarps = 4;
constexpr u32 kTileBytes = 512;
constexpr u32 kShortMax = 64;    // groups whose literal runs and matches are all shorter go through the tile
constexpr u32 kRunLit = 0x40000000u;  // `off` field of a literal run: tile offset - off is far below the tile
constexpr u32 kMark = 0xFFFFu;
constexpr u32 kExecRowBlock = 16;  // rows whose loads are in flight together

// Per-warp shared memory of k_seq_execute.
struct __align__(16) ExecWarpSmem {
  uint4 runs[72];              // run table (64 runs + slack for the lookups of lanes past the end): {delta lo, delta hi, match offset | kRunLit, -}
  u8 tile[kTileBytes + 128];
  u16 mk[kTileBytes + 128];    // run-start marks; then, for bytes with an in-tile source, the source's tile offset
};

// ------------------------------------------------------------------------------------------
// Whole-warp forward copy of n bytes between regions that do not overlap: 16-byte stores to the
// aligned body of dst, the source words funnel-shifted into place (src and dst may have any
// alignment). Reads whole aligned words, i.e. up to 3 bytes either side of the source range.
ZRA_DEV void warp_copy_wide(u8* dst, const u8* src, u32 n, u32 lane) {
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

// ---- lock-step short copies (direct path) -----------------------------------------------
// Every lane copies its own n bytes (possibly 0); m is a warp-uniform upper bound of n. Four bytes
// per trip, written as predicated PTX (one predicate per byte position, loads before stores,
// immediate offsets): nvcc turns the equivalent C++ into nested divergent branches.
#if defined(__CUDA_ARCH__)
#define ZRA_PRED4 "setp.gt.s32 p0, %2, 0;\n\tsetp.gt.s32 p1, %2, 1;\n\tsetp.gt.s32 p2, %2, 2;\n\tsetp.gt.s32 p3, %2, 3;\n\t"
#define ZRA_COPY4(LD, ST)                                                                               \
  "{\n\t.reg .pred p0, p1, p2, p3;\n\t.reg .b32 b0, b1, b2, b3;\n\t" ZRA_PRED4                          \
  "@p0 " LD " b0, [%0];\n\t@p1 " LD " b1, [%0+1];\n\t@p2 " LD " b2, [%0+2];\n\t@p3 " LD " b3, [%0+3];\n\t" \
  "@p0 " ST " [%1], b0;\n\t@p1 " ST " [%1+1], b1;\n\t@p2 " ST " [%1+2], b2;\n\t@p3 " ST " [%1+3], b3;\n\t}"
// global (read-only data: literal scratch / input) -> global
ZRA_DEV void lanes_copy_ro(u8* d, const u8* s, u32 n, u32 m) {
  i32 r = (i32)n;
#pragma unroll 1
  for (u32 k = 0; k < m; k += 4) {
    asm volatile(ZRA_COPY4("ld.global.nc.u8", "st.global.u8")::"l"(s), "l"(d), "r"(r) : "memory");
    s += 4; d += 4; r -= 4;
  }
}
// global (output written earlier by this warp) -> global
ZRA_DEV void lanes_copy_gg(u8* d, const u8* s, u32 n, u32 m) {
  i32 r = (i32)n;
#pragma unroll 1
  for (u32 k = 0; k < m; k += 4) {
    asm volatile(ZRA_COPY4("ld.global.u8", "st.global.u8")::"l"(s), "l"(d), "r"(r) : "memory");
    s += 4; d += 4; r -= 4;
  }
}
#else
ZRA_DEV void lanes_copy_ro(u8* d, const u8* s, u32 n, u32) { for (u32 i = 0; i < n; i++) d[i] = s[i]; }
ZRA_DEV void lanes_copy_gg(u8* d, const u8* s, u32 n, u32) { for (u32 i = 0; i < n; i++) d[i] = s[i]; }
#endif

ZRA_DEV u64 shfl64(u64 v, int srcLane) {
  u32 lo = __shfl_sync(kExecFull, (u32)v, srcLane), hi = __shfl_sync(kExecFull, (u32)(v >> 32), srcLane);
  return (u64)lo | ((u64)hi << 32);
}
ZRA_DEV u64 shfl64_up1(u64 v) {
  u32 lo = __shfl_up_sync(kExecFull, (u32)v, 1), hi = __shfl_up_sync(kExecFull, (u32)(v >> 32), 1);
  return (u64)lo | ((u64)hi << 32);
}
// byte load under a predicate (coherent: the source may be output this warp wrote earlier)
ZRA_DEV u32 ld_u8_if(const u8* p, bool on) {
#if defined(__CUDA_ARCH__)
  u32 v = 0;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.u8 %0, [%1];\n\t}" : "+r"(v) : "l"(p), "r"((u32)on) : "memory");
  return v;
#else
  return on ? *p : 0u;
#endif
}
ZRA_DEV u64 ld_rec(const u64* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// ------------------------------------------------------------------------------------------
// The current block of frame `warp`, executed by one warp. `sm` is the warp's shared memory; its mark array must be
// all zero on entry and is all zero again on return.
ZRA_DEV void exec_block(const u8* src, u8* dst, const FrameDesc& d, const FrameCtx& c, const u8* lit, u32 litStride, const u64* seqs,
                        u32 seqStride, u32 warp, u32 lane, ExecWarpSmem& sm) {
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
  u8* tile = sm.tile;
  const bool rle = c.litMode == LIT_RLE;
  const u8 rleByte = (u8)c.litSrc;
  const u8* litp = c.litMode == LIT_HUF ? lit + (u64)warp * litStride : fsrc + c.litSrc;
  const u64* sq = seqs + (u64)warp * seqStride;
  const u32 nbSeq = c.nbSeq;
  u8* blk = frame + blkDst;  // block-relative positions (the records' outEnd) index this
  const u32 laneLt = (1u << lane) - 1u, laneLe = laneLt | (1u << lane);
  u64 carry = 0;  // record of the last sequence of the previous iteration
  u32 pend = 0;   // bytes of the last tile group's partial final vector, kept in tile[0, pend) (warp-uniform)
  // lanes past the end repeat the last record (ll = ml = 0); the next group's records are requested a
  // whole iteration ahead
  u64 sNext = nbSeq ? ld_rec(sq + (lane < nbSeq ? lane : nbSeq - 1)) : 0ull;
  for (u32 base = 0; base < nbSeq; base += 32) {
    const u64 s = sNext;
    if (base + 32 < nbSeq) {
      const u32 nidx = base + 32 + lane;
      sNext = ld_rec(sq + (nidx < nbSeq ? nidx : nbSeq - 1));
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
    const bool viaTile = !rle && S <= kTileBytes - 16 && !__any_sync(kExecFull, ll >= kShortMax || ml >= kShortMax);
    if (viaTile) {
      // tile byte t <-> block byte S0 - a + t, a = 16-byte phase of the group's first output byte. When the previous
      // group went through the tile too, tile[0, a) holds its last, partial vector (a == pend).
      const u32 a = (u32)(reinterpret_cast<uintptr_t>(blk + S0) & 15u);
      u8* gbase = blk + ((i32)S0 - (i32)a);  // 16-byte aligned; tile byte t is gbase[t]
      const bool headValid = pend != 0;
      const u32 tl = po - S0 + a;  // tile offset of this lane's literals
      const u32 tm = tl + ll;      // ... and of its match
      // ---- run table + marks (runs are numbered in output order; empty literal runs do not count)
      {
        const u32 litMask = __ballot_sync(kExecFull, ll > 0);
        const u32 li = lane + (u32)__popc(litMask & laneLt);
        if (ll) {
          const i64 dl = (i64)(reinterpret_cast<intptr_t>(litp + pl) - reinterpret_cast<intptr_t>(gbase + tl));
          sm.runs[li] = make_uint4((u32)dl, (u32)((u64)dl >> 32), kRunLit, 0u);
          sm.mk[tl] = (u16)kMark;
        }
        if (ml) {
          const i64 dm = -(i64)off;
          sm.runs[li + (ll ? 1u : 0u)] = make_uint4((u32)dm, (u32)((u64)dm >> 32), off, 0u);
          sm.mk[tm] = (u16)kMark;
        }
      }
      __syncwarp();
      // ---- pass one: rows of 32 bytes; final sources are fetched, in-tile sources noted. Straight-line, predicated code:
      // the lanes are unrelated bytes, a branch would serialise them. Rows past the group's end (the batch of four is
      // not always full) have no marks and no active lane.
      const i32 end = (i32)(a + S);
      const u32 nrows = ((u32)end + 31u) >> 5;
      const uint4* runp = sm.runs - 1;  // entry of the last run started before the row
      u32 depRows = 0;                  // rows in which this lane has a byte with an in-tile source
      u32 bit = 1u;
      const u8* rowp = gbase + lane;    // where this lane's byte of the batch's first row goes in the output buffer
      u16* mkp = sm.mk + lane;
      u8* tp = tile + lane;
      i32 tb = (i32)lane;
      for (u32 r0 = 0; r0 < nrows; r0 += kExecRowBlock) {
#if defined(__CUDA_ARCH__)
        // A block of up to 16 rows in PTX (the C++ below, which the host logic build runs, is the specification). ALL the
        // block's byte loads are issued before the first store waits for one: the kernel is bound by the latency of these
        // loads (most match sources are far back in the frame and come from L2 or HBM), so a warp keeps a whole group in
        // flight instead of four rows (profiles/r03b: 31 % of the stall samples on the load, 17 % on the first store).
        // A value register holds 0xFFFFFFFF until its (predicated) load lands, which is the store's predicate.
        {
          const u32 mkS = (u32)__cvta_generic_to_shared(mkp), tpS = (u32)__cvta_generic_to_shared(tp);
          u32 runS = (u32)__cvta_generic_to_shared(runp);
          const u32 runS0 = runS;
          const u32 left = nrows - r0;
          asm volatile(
              "{\n\t.reg .pred pm, pa, pact, pdep, pst, pu;\n\t"
              ".reg .b32 m, mask, x, c, ea, e0, e1, e2, e3, t, ts, bj;\n\t.reg .b64 d, sp;\n\t"
              ".reg .b32 v0, v1, v2, v3, v4, v5, v6, v7, v8, v9, v10, v11, v12, v13, v14, v15;\n\t"
#define ZRA_EXEC_ROW(MKOFF, TOFF, J, V, LBL)                                                                       \
              "mov.b32 " V ", 0xffffffff;\n\tsetp.le.u32 pu, %10, " J ";\n\t@pu bra.uni LROWS_DONE;\n\t"                 \
              "ld.shared.u16 m, [%2+" MKOFF "];\n\tsetp.ne.u32 pm, m, 0;\n\tvote.sync.ballot.b32 mask, pm, 0xffffffff;\n\t" \
              "and.b32 x, mask, %7;\n\tpopc.b32 c, x;\n\tmad.lo.u32 ea, c, 16, %0;\n\tld.shared.v4.u32 {e0, e1, e2, e3}, [ea];\n\t"    \
              "popc.b32 c, mask;\n\tmad.lo.u32 %0, c, 16, %0;\n\t"                                             \
              "add.s32 t, %4, " TOFF ";\n\tsub.s32 ts, t, e2;\n\t"                                             \
              "setp.ge.s32 pa, t, %5;\n\tsetp.lt.and.s32 pact, t, %6, pa;\n\t"                                 \
              "setp.ge.and.s32 pdep, ts, %5, pact;\n\tsetp.lt.and.s32 pst, ts, %5, pact;\n\t"                  \
              "@pdep st.shared.u16 [%2+" MKOFF "], ts;\n\tshl.b32 bj, %8, " J ";\n\t@pdep or.b32 %1, %1, bj;\n\t" \
              "mov.b64 d, {e0, e1};\n\tadd.s64 sp, %3, d;\n\t@pst ld.global.u8 " V ", [sp+" TOFF "];\n\t"
              ZRA_EXEC_ROW("0", "0", "0", "v0", "L0")
              ZRA_EXEC_ROW("64", "32", "1", "v1", "L1")
              ZRA_EXEC_ROW("128", "64", "2", "v2", "L2")
              ZRA_EXEC_ROW("192", "96", "3", "v3", "L3")
              ZRA_EXEC_ROW("256", "128", "4", "v4", "L4")
              ZRA_EXEC_ROW("320", "160", "5", "v5", "L5")
              ZRA_EXEC_ROW("384", "192", "6", "v6", "L6")
              ZRA_EXEC_ROW("448", "224", "7", "v7", "L7")
              ZRA_EXEC_ROW("512", "256", "8", "v8", "L8")
              ZRA_EXEC_ROW("576", "288", "9", "v9", "L9")
              ZRA_EXEC_ROW("640", "320", "10", "v10", "L10")
              ZRA_EXEC_ROW("704", "352", "11", "v11", "L11")
              ZRA_EXEC_ROW("768", "384", "12", "v12", "L12")
              ZRA_EXEC_ROW("832", "416", "13", "v13", "L13")
              ZRA_EXEC_ROW("896", "448", "14", "v14", "L14")
              ZRA_EXEC_ROW("960", "480", "15", "v15", "L15")
#undef ZRA_EXEC_ROW
              "LROWS_DONE:\n\t"
#define ZRA_EXEC_ST(TOFF, V) "setp.lt.u32 pst, " V ", 256;\n\t@pst st.shared.u8 [%9+" TOFF "], " V ";\n\t"
              ZRA_EXEC_ST("0", "v0")
              ZRA_EXEC_ST("32", "v1")
              ZRA_EXEC_ST("64", "v2")
              ZRA_EXEC_ST("96", "v3")
              ZRA_EXEC_ST("128", "v4")
              ZRA_EXEC_ST("160", "v5")
              ZRA_EXEC_ST("192", "v6")
              ZRA_EXEC_ST("224", "v7")
              ZRA_EXEC_ST("256", "v8")
              ZRA_EXEC_ST("288", "v9")
              ZRA_EXEC_ST("320", "v10")
              ZRA_EXEC_ST("352", "v11")
              ZRA_EXEC_ST("384", "v12")
              ZRA_EXEC_ST("416", "v13")
              ZRA_EXEC_ST("448", "v14")
              ZRA_EXEC_ST("480", "v15")
#undef ZRA_EXEC_ST
              "}"
              : "+r"(runS), "+r"(depRows)
              : "r"(mkS), "l"(rowp), "r"(tb), "r"((i32)a), "r"(end), "r"(laneLe), "r"(bit), "r"(tpS), "r"(left)
              : "memory");
          runp = reinterpret_cast<const uint4*>(reinterpret_cast<const u8*>(runp) + (runS - runS0));
        }
#else
        u32 val[kExecRowBlock];
        bool st[kExecRowBlock];
        for (u32 j = 0; j < kExecRowBlock; j++) {
          st[j] = false; val[j] = 0;
          if (r0 + j >= nrows) break;
          const u32 mask = __ballot_sync(kExecFull, mkp[32 * j] != 0);
          const uint4 e = runp[__popc(mask & laneLe)];
          runp += __popc(mask);
          const i32 t = tb + 32 * (i32)j;
          const i32 ts = t - (i32)e.z;
          const bool act = t >= (i32)a && t < end;
          const bool dep = act && ts >= (i32)a;
          st[j] = act && ts < (i32)a;
          if (dep) { mkp[32 * j] = (u16)ts; depRows |= bit << j; }
          val[j] = ld_u8_if(rowp + 32 * j + (i64)((u64)e.x | ((u64)e.y << 32)), st[j]);
        }
        for (u32 j = 0; j < kExecRowBlock; j++)
          if (st[j]) tp[32 * j] = (u8)val[j];
#endif
        rowp += 32 * kExecRowBlock; mkp += 32 * kExecRowBlock; tp += 32 * kExecRowBlock; tb += 32 * kExecRowBlock; bit <<= kExecRowBlock;
      }
      __syncwarp();
      // ---- pass two: bytes whose source is in the tile, row by row
      u32 rows = __reduce_or_sync(kExecFull, depRows);
      while (rows) {
        const u32 r = (u32)__ffs((int)rows) - 1u;
        rows &= rows - 1u;
        const u32 t = 32u * r + lane;
        const bool mine = (depRows >> r) & 1u;
        u32 ts = 0;
        if (mine) { ts = sm.mk[t]; sm.mk[t] = 0; }
        bool waiting = mine && ts >= 32u * r;  // depends on a byte of its own row
        if (mine && !waiting) tile[t] = tile[ts];
        __syncwarp();
        if (__any_sync(kExecFull, waiting)) {
          for (;;) {
            const u32 wm = __ballot_sync(kExecFull, waiting);
            if (!wm) break;
            // a source always precedes its destination: the lowest waiting lane is always ready
            if (waiting && !((wm >> (ts & 31u)) & 1u)) { tile[t] = tile[ts]; waiting = false; }
            __syncwarp();
          }
        }
      }
      // the marks go (a mark that became a source note was cleared above)
      if (ll) sm.mk[tl] = 0;
      if (ml) sm.mk[tm] = 0;
      // ---- the group leaves: whole 16-byte vectors [firstFull, endA); what is left either side goes by bytes
      {
        const u32 endB = (u32)end, endA = endB & ~15u;
        const u32 firstFull = headValid ? 0u : (a + 15u) & ~15u;
        for (u32 lo = firstFull + 16u * lane; lo < endA; lo += 512u)
          *reinterpret_cast<uint4*>(gbase + lo) = *reinterpret_cast<const uint4*>(tile + lo);
        if (!headValid && a) {  // ragged start (first tile group of a block, or after a direct group)
          const u32 i = a + lane, stop = firstFull < endB ? firstFull : endB;
          if (i < stop) gbase[i] = tile[i];
        }
        const u32 tailLo = endA > firstFull ? endA : firstFull;  // bytes past the last whole vector
        if (tailLo + lane < endB) gbase[tailLo + lane] = tile[tailLo + lane];
        // ... and they are carried, so that the next group's first vector is whole
        const u32 newPend = (endB > endA && endA >= firstFull) ? endB - endA : 0u;
        u8 keep = 0;
        if (lane < newPend) keep = tile[endA + lane];
        __syncwarp();
        if (lane < newPend) tile[lane] = keep;
        pend = newPend;
      }
      __syncwarp();
      continue;
    }
    // ---- direct path (long runs / matches): straight to the output buffer
    pend = 0;
    const u32 myDst = blkDst + po;
    u32 longLit = __ballot_sync(kExecFull, ll >= kLongCopy);
    while (longLit) {
      int who = __ffs((int)longLit) - 1;
      longLit &= longLit - 1;
      u32 L = __shfl_sync(kExecFull, ll, who), from = __shfl_sync(kExecFull, pl, who), to = __shfl_sync(kExecFull, myDst, who);
      if (rle) { for (u32 i = lane; i < L; i += 32) frame[to + i] = rleByte; }
      else warp_copy_wide(frame + to, litp + from, L, lane);
    }
    {
      const u32 n = ll < kLongCopy ? ll : 0;
      const u32 m = __reduce_max_sync(kExecFull, n);
      if (rle) { for (u32 i = 0; i < n; i++) frame[myDst + i] = rleByte; }
      else lanes_copy_ro(frame + myDst, litp + pl, n, m);
    }
    __syncwarp();
    const u32 mpos = myDst + ll;
    const u32 msrc = mpos - off;
    bool pending = ml > 0;
    for (;;) {
      u32 mask = __ballot_sync(kExecFull, pending);
      if (!mask) break;
      int first = __ffs((int)mask) - 1;
      u32 hwm = __shfl_sync(kExecFull, mpos, first);
      u32 fml = __shfl_sync(kExecFull, ml, first);
      if (fml >= kLongCopy) {
        u32 fs = __shfl_sync(kExecFull, msrc, first), fo = __shfl_sync(kExecFull, off, first);
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
        const u32 m = __reduce_max_sync(kExecFull, n);
        lanes_copy_gg(frame + mpos, frame + msrc, n, m);
        if (ready) pending = false;
      }
      __syncwarp();
    }
  }
  // trailing literals
  const u32 litPos = rec_lit_end(carry);
  const u32 pos = blkDst + rec_out_end(carry);
  const u32 rest = c.litSize - litPos;
  if (rle) { for (u32 i = lane; i < rest; i += 32) frame[pos + i] = rleByte; }
  else warp_copy_wide(frame + pos, litp + litPos, rest, lane);
}

// ------------------------------------------------------------------------------------------
// Sequence execution: one warp per frame (exec_block above), kExecWarps frames per CTA.
__global__ void __launch_bounds__(kExecWarps * 32) k_seq_execute(const u8* __restrict__ src, u8* dst, const FrameDesc* __restrict__ descs,
                                                                  const FrameCtx* __restrict__ ctxs, const u8* __restrict__ lit, u32 litStride,
                                                                  const u64* __restrict__ seqs, u32 seqStride, u32 nFrames) {
  __shared__ ExecWarpSmem sm[kExecWarps];
  const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  if (warp >= nFrames) return;
  const FrameCtx& c = ctxs[warp];
  if (c.status || c.blkType == BT_NONE) return;
  ExecWarpSmem& mine = sm[threadIdx.x >> 5];
  if (c.blkType == BT_COMPRESSED) {
    for (u32 i = lane; i < (kTileBytes + 128) / 2; i += 32) reinterpret_cast<u32*>(mine.mk)[i] = 0u;
    __syncwarp();
  }
  exec_block(src, dst, descs[warp], c, lit, litStride, seqs, seqStride, warp, lane, mine);
}

}  // namespace zrab
