// encode_kernels.cu — sm_100a kernels of the frame-parallel zstd encoder (levels 1-3) and launcher.
//
// One round encodes block r of every frame of the batch (frames <= 128 KiB need one round):
//   k_enc_begin     1 thread / frame   frame state, parameters
//   k_enc_match     1 thread / frame   greedy hash match finder (per-frame tables in HBM)
//   k_enc_literals  1 thread / frame   literal gather + byte histogram
//   k_enc_plan      1 thread / frame   Huffman code, FSE tables, section headers
//   k_enc_huf       1 thread / stream  Huffman streams (4 per block)
//   k_enc_seq       1 thread / frame   FSE sequence bitstream
//   k_enc_assemble  1 warp   / frame   block header + sections into the frame's slot
//   k_enc_finish    4 threads/ frame   XXH64 content checksum, final frame size
// Replaces the per-frame ZSTD_compress2 calls of zra::CompressBuffer / Compressor::Compress
// (source/zra.cpp:216-225, 329-338). The thread bodies are in enc_core.cuh.
#include <cuda_runtime.h>

#include <cstdlib>

#include "enc_core.cuh"
#include "encode_launch.h"

namespace zrab {

namespace {
constexpr u32 kFullMask = 0xFFFFFFFFu;

struct EncView {
  u8* s;
  EncodeLayout lay;
  __device__ EncCtx& ctx(u32 i) const { return reinterpret_cast<EncCtx*>(s + lay.offCtx)[i]; }
  __device__ EncScratch frame(u32 i) const {
    EncScratch f;
    f.tabS = reinterpret_cast<u32*>(s + lay.offTabS) + (u64)i * lay.tabSEntries;
    f.tabL = reinterpret_cast<u32*>(s + lay.offTabL) + (u64)i * lay.tabLEntries;
    f.seqs = reinterpret_cast<u64*>(s + lay.offSeqs) + (u64)i * lay.seqStride;
    f.lit = s + lay.offLit + (u64)i * lay.litStride;
    f.hist = reinterpret_cast<u32*>(s + lay.offHist) + (u64)i * 256;
    f.hcodes = reinterpret_cast<HufCode*>(s + lay.offCodes) + (u64)i * 256;
    f.hufOut = s + lay.offHuf + (u64)i * 4 * lay.hufStride;
    f.hufStride = lay.hufStride;
    f.hdr = s + lay.offHdr + (u64)i * 512;
    f.tt = reinterpret_cast<FseSymTT*>(s + lay.offTT) + (u64)i * 128;
    f.states = reinterpret_cast<u16*>(s + lay.offStates) + (u64)i * 1280;
    f.seqOut = s + lay.offSeqOut + (u64)i * lay.seqOutStride;
    f.seqOutCap = lay.seqOutStride;
    f.cells = s + lay.offCells + (u64)i * 1024;
    f.cnt = (lay.ctaMatch || lay.ctaBig) ? reinterpret_cast<u32*>(s + lay.offCnt) + (u64)i * 128 : nullptr;
    return f;
  }
  __device__ u8* out(u32 i) const { return s + lay.offOut + (u64)i * lay.outStride; }
};

struct EncJob {
  const u8* in;
  u64 inOff, inEnd;
  u32 frameSize, nFrames;
  int level;
  u32 checksum;
};

__device__ __forceinline__ u32 frame_len(const EncJob& j, u32 i) {
  u64 begin = j.inOff + (u64)i * j.frameSize;
  u64 left = j.inEnd - begin;
  return (u32)(left < j.frameSize ? left : j.frameSize);
}
}  // namespace

__global__ void k_enc_begin(EncJob j, EncView v) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.nFrames) return;
  EncCtx& c = v.ctx(i);
  u32 len = frame_len(j, i);
  EncParams p = enc_params(j.level, len, j.checksum != 0);
  c.status = 0;
  c.srcLen = len;
  c.rep[0] = 1; c.rep[1] = 4; c.rep[2] = 8;
  c.outPos = 0;
  c.windowLog = p.windowLogMax;
  c.blkActive = 0;
}

__global__ void k_enc_match(EncJob j, EncView v, u32 round) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.nFrames) return;
  EncCtx c = v.ctx(i);
  u32 pos = round * kBlockSizeMax;
  c.blkActive = pos < c.srcLen || (round == 0);
  if (c.blkActive) {
    c.blkPos = pos;
    c.blkLen = c.srcLen - pos < kBlockSizeMax ? c.srcLen - pos : kBlockSizeMax;
    c.lastBlock = pos + c.blkLen >= c.srcLen;
    EncParams p = enc_params(j.level, c.srcLen, j.checksum != 0);
    enc_match(j.in, j.inOff + (u64)i * j.frameSize, p, c, v.frame(i));
  }
  v.ctx(i) = c;
}

// ------------------------------------------------------------------------------------------
// Frame-cooperative match finder (frames <= 64 KiB: one block, positions fit 16 bits).
// One CTA per frame; the frame's hash tables live in shared memory as 16-bit positions (level 3,
// 64 KiB frames: 2^15 short + 2^16 long entries = 192 KiB, i.e. one frame per SM). The block is
// consumed in rounds of T = blockDim positions:
//   A  every thread hashes its position (8-byte long hash, mls-byte short hash), reads the tables,
//      inserts itself, and after a barrier reads again: the newest candidate BELOW its own position
//      wins (all positions are inserted, unlike the serial reference matcher, which only inserts
//      the positions it visits — the parse is at least as dense);
//   B  every thread verifies its candidates against the input (L1/L2) and leaves the best
//      (length <= kLaneLenCap, offset) in shared memory;
//   C  warp 0 walks the round greedily with warp-uniform state: next position with a match by
//      ballot/ffs, one-byte lazy step towards a long match, warp-wide extension of capped
//      matches, repeat-offset coding, sequence record. Nothing in the walk waits for a load.
// After the last round the whole CTA gathers the literals (one thread per sequence) and builds the
// literal histogram and the LL / OF / ML code histograms in shared memory.
// Replaces ZSTD_compressBlock_fast / _doubleFast for these frames (zstd_fast.c:46-183,
// zstd_double_fast.c:50-316): same hash functions and minimum match lengths; a different
// (parallel) visiting order, so the bytes differ from the reference's while the format, the
// decoder and the ratio (tests: within 3 %) do not.
constexpr u32 kLaneLenCap = 40;
constexpr u32 kInfoMore = 0x80000000u;  // the lane stopped comparing before the first mismatch

__device__ __forceinline__ u64 gld8(const u8* base, u64 off) {
  const u32* w = reinterpret_cast<const u32*>(base) + (off >> 2);
  const u32 sh = (u32)(off & 3) * 8;
  const u32 a = __ldg(w), b = __ldg(w + 1), c = sh ? __ldg(w + 2) : 0u;
  return (u64)__funnelshift_r(a, b, sh) | ((u64)__funnelshift_r(b, c, sh) << 32);
}
__device__ __forceinline__ u32 gld4(const u8* base, u64 off) {
  const u32* w = reinterpret_cast<const u32*>(base) + (off >> 2);
  const u32 sh = (u32)(off & 3) * 8;
  const u32 a = __ldg(w), b = sh ? __ldg(w + 1) : 0u;
  return __funnelshift_r(a, b, sh);
}
__device__ __forceinline__ u32 common8(u64 x) { return x ? ((u32)__ffsll((long long)x) - 1) >> 3 : 8u; }

// Sequence codes (zstd_compress_internal.h: ZSTD_LLcode / ZSTD_MLcode), table-free.
__device__ __forceinline__ u32 ll_code_fast(u32 ll) {
  if (ll < 16) return ll;
  if (ll >= 64) return highbit32(ll) + 19;
  if (ll < 24) return 16 + ((ll - 16) >> 1);
  if (ll < 32) return 20 + ((ll - 24) >> 2);
  if (ll < 48) return 22 + ((ll - 32) >> 3);
  return 24;
}
__device__ __forceinline__ u32 ml_code_fast(u32 mlBase) {  // mlBase = matchLength - 3
  if (mlBase < 32) return mlBase;
  if (mlBase >= 128) return highbit32(mlBase) + 36;
  if (mlBase < 40) return 32 + ((mlBase - 32) >> 1);   // 35,37,39,41 -> codes 32..35
  if (mlBase < 48) return 36 + ((mlBase - 40) >> 2);   // 43,47 -> 36,37
  if (mlBase < 64) return 38 + ((mlBase - 48) >> 3);   // 51,59 -> 38,39
  if (mlBase < 96) return 40 + ((mlBase - 64) >> 4);   // 67,83 -> 40,41
  return 42;                                           // 99..130
}

__device__ __forceinline__ void table_insert_min(u16* cell, u32 pos, u32 roundBase) {
  unsigned short cur = *cell;
  while ((u16)(cur - roundBase) > (u16)(pos - roundBase)) {
    const unsigned short prev = atomicCAS(reinterpret_cast<unsigned short*>(cell), cur, (unsigned short)pos);
    if (prev == cur) break;
    cur = prev;
  }
}

template <int T>
__global__ void __launch_bounds__(T) k_enc_match_cta(EncJob j, EncView v) {
  extern __shared__ __align__(16) u8 smem[];
  __shared__ u32 sAnchor;
  __shared__ u32 sFinal[8];
  const u32 i = blockIdx.x;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  EncCtx& gc = v.ctx(i);
  const u32 len = gc.srcLen;
  const u32 logS = v.lay.matchLogS, logL = v.lay.matchLogL, mls = v.lay.matchMls;
  const bool dfast = logL != 0;
  const u32 nS = 1u << logS, nL = dfast ? (1u << logL) : 0u;
  u16* tabS = reinterpret_cast<u16*>(smem);
  u16* tabL = tabS + nS;
  u32* info = reinterpret_cast<u32*>(smem + 2ull * (nS + nL));
  u32* hist = info + T;   // 256 literal counts, then 36 + 32 + 53 code counts
  u32* cnt = hist + 256;
  u16* take = reinterpret_cast<u16*>(cnt + 128);  // per position of the round: the match a walk arriving there takes
  __shared__ u32 sHas[T / 32];                     // per 32 positions: which of them have a match
  const u8* base = j.in;
  const u64 fbase = j.inOff + (u64)i * j.frameSize;
  const EncScratch s = v.frame(i);
  u32* side = reinterpret_cast<u32*>(s.hufOut);  // per sequence: literal source | literal destination << 16
  // ---- clear tables and histograms
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const u32 n16 = (2u * (nS + nL)) >> 4;
    for (u32 k = tid; k < n16; k += T) z[k] = make_uint4(0, 0, 0, 0);
    for (u32 k = tid; k < 256 + 128; k += T) hist[k] = 0;
    if (tid == 0) sAnchor = 0;
  }
  __syncthreads();
  // selector state (meaningful in warp 0, warp-uniform)
  EncCtx rc;
  rc.rep[0] = 1; rc.rep[1] = 4; rc.rep[2] = 8;
  u32 anchor = 0, nseq = 0, litPos = 0;
  const u32 hashEnd = len >= 16 ? len - 8 : 0;  // positions below this can start a match (zstd: ip < iend - 8)
  for (u32 rb = 0; rb < hashEnd; rb += T) {
    const u32 pos = rb + tid;
    const u32 curAnchor = sAnchor;
    if (curAnchor >= rb + T) continue;  // the whole round lies inside a match already taken (uniform)
    // ---- A: hash, read the pre-round candidates, then insert. The insert keeps the LOWEST position
    // of this round per cell (16-bit compare-and-swap on the distance from the round base, under
    // which every older entry ranks above the round's own), so the table — and with it the archive —
    // does not depend on thread timing, and later positions of the round see an in-round candidate.
    const bool live = pos < hashEnd;
    u64 x = 0;
    u32 hS = 0, hL = 0, cS = 0, cL = 0;
    if (live) {
      x = gld8(base, fbase + pos);
      hS = mls <= 4 ? ((u32)x * 2654435761u) >> (32 - logS) : (u32)(((x << (64 - 8 * mls)) * 0x9E3779B97F4A7C15ull) >> (64 - logS));
      cS = tabS[hS];
      if (dfast) {
        hL = (u32)((x * 0xCF1BBCDCB7A56463ull) >> (64 - logL));
        cL = tabL[hL];
      }
    }
    __syncthreads();
    if (live) {
      table_insert_min(&tabS[hS], pos, rb);
      if (dfast) table_insert_min(&tabL[hL], pos, rb);
    }
    __syncthreads();
    u32 e = 0;
    if (live && pos >= curAnchor && pos > 0) {
      u32 c2 = tabS[hS];
      if (c2 < pos && c2 >= rb) cS = c2;
      bool okS = cS < pos, okL = false;
      if (dfast) {
        c2 = tabL[hL];
        if (c2 < pos && c2 >= rb) cL = c2;
        okL = cL < pos;
      }
      // ---- B: verify
      u32 bestLen = 0, bestOff = 0;
      u32 nL8 = 0, nS8 = 0;
      if (okL) nL8 = common8(gld8(base, fbase + cL) ^ x);
      if (okS && (!okL || cS != cL)) nS8 = common8(gld8(base, fbase + cS) ^ x);
      if (nL8 >= 4 && nL8 >= nS8) { bestLen = nL8; bestOff = pos - cL; }
      else if (nS8 >= 4) { bestLen = nS8; bestOff = pos - cS; }
      bool more = false;
      if (bestLen == 8) {
        const u32 cand = pos - bestOff;
        more = true;
        while (bestLen < kLaneLenCap && pos + bestLen + 8 <= len) {
          const u32 c = common8(gld8(base, fbase + pos + bestLen) ^ gld8(base, fbase + cand + bestLen));
          bestLen += c;
          if (c < 8) { more = false; break; }
        }
      }
      if (bestLen >= 4) e = bestOff | (bestLen << 16) | (more ? kInfoMore : 0u);
    }
    info[tid] = e;
    {
      const u32 bal = __ballot_sync(kFullMask, e != 0);
      if (lane == 0) sHas[warp] = bal;
    }
    __syncthreads();
    // ---- B2 (all threads): which match does a walk take that ARRIVES at position rb + tid? The next position at or
    // after it that has a match, moved one byte on when that match is short (< 8) and its right neighbour's is not
    // (the one-byte lazy step; never across a 32-position group, as in the selection loop this replaces). Doing the
    // search here, in parallel, leaves the serial walk two shared-memory loads per sequence.
    {
      u32 g = warp;
      u32 w = sHas[g] & (0xFFFFFFFFu << lane);
      while (!w && ++g < T / 32) w = sHas[g];
      u32 q = 0xFFFFu;
      if (w) {
        q = g * 32 + ((u32)__ffs((int)w) - 1);
        if ((q & 31u) != 31u && ((sHas[q >> 5] >> ((q & 31u) + 1u)) & 1u)) {
          const u32 l1 = (info[q] >> 16) & 0xFFu, l2 = (info[q + 1] >> 16) & 0xFFu;
          if (l1 < 8 && l2 >= 8) q++;
        }
      }
      take[tid] = (u16)q;
    }
    __syncthreads();
    // ---- C: greedy walk by warp 0 (warp-uniform state; the lanes only differ in the warp-wide extension)
    if (warp == 0) {
      {
        for (;;) {
          const u32 arrive = anchor > rb ? anchor - rb : 0;
          if (arrive >= T) break;
          const u32 k = take[arrive];
          if (k == 0xFFFFu) break;
          const u32 ee = info[k];
          const u32 gb = rb;
          const u32 mpos = gb + k;
          u32 ml = (ee >> 16) & 0xFFu;
          const u32 off = ee & 0xFFFFu;
          if (ee & kInfoMore) {
            // warp-wide extension: 4 bytes per lane and trip
            for (;;) {
              const u32 a = mpos + ml + 4 * lane;
              u32 diff = 0;  // byte k of diff != 0: byte k differs / is past the end
              if (a + 4 <= len) {
                diff = gld4(base, fbase + a) ^ gld4(base, fbase + a - off);
              } else {
                for (u32 b = 0; b < 4; b++) {
                  if (a + b >= len || base[fbase + a + b] != base[fbase + a + b - off]) { diff = 0xFFu << (8 * b); break; }
                }
              }
              const u32 bad = __ballot_sync(kFullMask, diff != 0);
              if (!bad) { ml += 128; continue; }
              const u32 first = (u32)__ffs((int)bad) - 1;
              const u32 d = __shfl_sync(kFullMask, diff, first);
              ml += 4 * first + (((u32)__ffs((int)d) - 1) >> 3);
              break;
            }
          }
          // ---- emit
          const u32 ll = mpos - anchor;
          const u64 rec = emit_sequence(rc, ll, ml, off);
          if (lane == 0) {
            s.seqs[nseq] = rec;
            side[nseq] = anchor | (litPos << 16);
          }
          nseq++;
          litPos += ll;
          anchor = mpos + ml;
        }
      }
      if (lane == 0) sAnchor = anchor;
    }
    __syncthreads();
  }
  // ---- literal gather + histograms (whole CTA), results
  if (tid == 0) {
    sFinal[0] = litPos; sFinal[1] = nseq; sFinal[2] = rc.rep[0]; sFinal[3] = rc.rep[1]; sFinal[4] = rc.rep[2];
    __threadfence_block();
  }
  __syncthreads();
  anchor = sAnchor;
  litPos = sFinal[0];
  nseq = sFinal[1];
  __threadfence();  // the records / side entries were written by warp 0 through global memory
  for (u32 q = tid; q < nseq; q += T) {
    const u64 rec = s.seqs[q];
    const u32 sd = side[q];
    const u32 ll = seq_ll(rec), from = sd & 0xFFFFu, to = sd >> 16;
    for (u32 k = 0; k < ll; k++) {
      const u8 b = base[fbase + from + k];
      s.lit[to + k] = b;
      atomicAdd(&hist[b], 1u);
    }
    atomicAdd(&cnt[ll_code_fast(ll)], 1u);
    atomicAdd(&cnt[36 + highbit32(seq_off(rec))], 1u);
    atomicAdd(&cnt[68 + ml_code_fast(seq_ml(rec) - 3)], 1u);
  }
  const u32 rest = len - anchor;
  for (u32 q = tid; q < rest; q += T) {
    const u8 b = base[fbase + anchor + q];
    s.lit[litPos + q] = b;
    atomicAdd(&hist[b], 1u);
  }
  __syncthreads();
  for (u32 k = tid; k < 256; k += T) s.hist[k] = hist[k];
  for (u32 k = tid; k < 128; k += T) s.cnt[k] = cnt[k];
  if (tid == 0) {
    gc.blkActive = 1;
    gc.blkPos = 0;
    gc.blkLen = len;
    gc.lastBlock = 1;
    gc.repSave[0] = 1; gc.repSave[1] = 4; gc.repSave[2] = 8;
    gc.rep[0] = sFinal[2]; gc.rep[1] = sFinal[3]; gc.rep[2] = sFinal[4];
    gc.nbSeq = nseq;
    gc.litSize = litPos + rest;
  }
}

// ------------------------------------------------------------------------------------------
// The same frame-cooperative matcher for frames ABOVE 64 KiB (one launch per 128 KiB block of every frame, `round`).
// Positions are frame-relative 32-bit numbers, the tables hold 32-bit entries and live in shared memory for the
// duration of a block: they are loaded from / saved to the frame's HBM table slots between blocks (a block can match into
// every earlier block of its frame — the window is the frame), and start empty at block 0. Match info per position is
// offset (24 bits) | length << 24 (7 bits) | "more" (bit 31): frames up to 16 MiB (larger ones keep the thread-per-frame
// matcher). The literal gather needs no per-sequence side array: sequence chunks are scanned block-wide for their literal
// and input offsets. Everything else — hashes, lowest-position-wins inserts, verification, the parallel "take" search,
// the warp-0 greedy walk with the one-byte lazy step and the warp-wide extension — is k_enc_match_cta's.
// Replaces the thread-per-frame k_enc_match + k_enc_literals for these frames (1.9 GB/s at 256 KiB frames, profiles/r03t).
constexpr u32 kBigMaxFrame = 1u << 24;

__device__ __forceinline__ void table_insert_min32(u32* cell, u32 pos, u32 roundBase) {
  u32 cur = *cell;
  while (cur - roundBase > pos - roundBase) {
    const u32 prev = atomicCAS(cell, cur, pos);
    if (prev == cur) break;
    cur = prev;
  }
}

template <int T>
__global__ void __launch_bounds__(T) k_enc_match_cta_big(EncJob j, EncView v, u32 round) {
  extern __shared__ __align__(16) u8 smem[];
  __shared__ u32 sAnchor;
  __shared__ u32 sFinal[8];
  __shared__ u32 sHas[T / 32];
  __shared__ u32 sScanL[T], sScanO[T];
  const u32 i = blockIdx.x;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  EncCtx& gc = v.ctx(i);
  const u32 flen = gc.srcLen;
  const u32 bstart = round * kBlockSizeMax;
  const bool activeBlk = bstart < flen || round == 0;
  if (!activeBlk) {
    if (tid == 0) gc.blkActive = 0;
    return;
  }
  const u32 blen = flen - bstart < kBlockSizeMax ? flen - bstart : kBlockSizeMax;
  const u32 bend = bstart + blen;
  const u32 logS = v.lay.matchLogS, logL = v.lay.matchLogL, mls = v.lay.matchMls;
  const bool dfast = logL != 0;
  const u32 nS = 1u << logS, nL = dfast ? (1u << logL) : 0u;
  u32* tabS = reinterpret_cast<u32*>(smem);
  u32* tabL = tabS + nS;
  u32* info = tabL + nL;
  u32* hist = info + T;   // 256 literal counts, then 36 + 32 + 53 code counts
  u32* cnt = hist + 256;
  u16* take = reinterpret_cast<u16*>(cnt + 128);
  const u8* base = j.in;
  const u64 fbase = j.inOff + (u64)i * j.frameSize;
  const EncScratch s = v.frame(i);
  // ---- tables: empty at block 0, else what the previous block of this frame left
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const u32 nSv = nS >> 2, nLv = nL >> 2;
    if (round == 0) {
      for (u32 k = tid; k < nSv + nLv; k += T) z[k] = make_uint4(0, 0, 0, 0);
    } else {
      const uint4* gS = reinterpret_cast<const uint4*>(s.tabS);
      const uint4* gL = reinterpret_cast<const uint4*>(s.tabL);
      for (u32 k = tid; k < nSv; k += T) z[k] = gS[k];
      for (u32 k = tid; k < nLv; k += T) z[nSv + k] = gL[k];
    }
    for (u32 k = tid; k < 256 + 128; k += T) hist[k] = 0;
    if (tid == 0) sAnchor = bstart;
  }
  __syncthreads();
  EncCtx rc;
  rc.rep[0] = gc.rep[0]; rc.rep[1] = gc.rep[1]; rc.rep[2] = gc.rep[2];
  const u32 rep0In = rc.rep[0], rep1In = rc.rep[1], rep2In = rc.rep[2];
  u32 anchor = bstart, nseq = 0, litPos = 0;
  const u32 hashEnd = blen >= 16 ? bend - 8 : bstart;  // positions below this can start a match
  for (u32 rb = bstart; rb < hashEnd; rb += T) {
    const u32 pos = rb + tid;
    const u32 curAnchor = sAnchor;
    if (curAnchor >= rb + T) continue;  // the whole round lies inside a match already taken (uniform)
    const bool live = pos < hashEnd;
    u64 x = 0;
    u32 hS = 0, hL = 0, cS = 0, cL = 0;
    if (live) {
      x = gld8(base, fbase + pos);
      hS = mls <= 4 ? ((u32)x * 2654435761u) >> (32 - logS) : (u32)(((x << (64 - 8 * mls)) * 0x9E3779B97F4A7C15ull) >> (64 - logS));
      cS = tabS[hS];
      if (dfast) {
        hL = (u32)((x * 0xCF1BBCDCB7A56463ull) >> (64 - logL));
        cL = tabL[hL];
      }
    }
    __syncthreads();
    if (live) {
      table_insert_min32(&tabS[hS], pos, rb);
      if (dfast) table_insert_min32(&tabL[hL], pos, rb);
    }
    __syncthreads();
    u32 e = 0;
    if (live && pos >= curAnchor && pos > 0) {
      u32 c2 = tabS[hS];
      if (c2 < pos && c2 >= rb) cS = c2;
      bool okS = cS < pos, okL = false;
      if (dfast) {
        c2 = tabL[hL];
        if (c2 < pos && c2 >= rb) cL = c2;
        okL = cL < pos;
      }
      u32 bestLen = 0, bestOff = 0;
      u32 nL8 = 0, nS8 = 0;
      if (okL) nL8 = common8(gld8(base, fbase + cL) ^ x);
      if (okS && (!okL || cS != cL)) nS8 = common8(gld8(base, fbase + cS) ^ x);
      if (nL8 >= 4 && nL8 >= nS8) { bestLen = nL8; bestOff = pos - cL; }
      else if (nS8 >= 4) { bestLen = nS8; bestOff = pos - cS; }
      bool more = false;
      if (bestLen == 8) {
        const u32 cand = pos - bestOff;
        more = true;
        while (bestLen < kLaneLenCap && pos + bestLen + 8 <= bend) {
          const u32 c = common8(gld8(base, fbase + pos + bestLen) ^ gld8(base, fbase + cand + bestLen));
          bestLen += c;
          if (c < 8) { more = false; break; }
        }
      }
      if (bestLen >= 4 && bestOff < kBigMaxFrame) e = bestOff | (bestLen << 24) | (more ? kInfoMore : 0u);
    }
    info[tid] = e;
    {
      const u32 bal = __ballot_sync(kFullMask, e != 0);
      if (lane == 0) sHas[warp] = bal;
    }
    __syncthreads();
    {
      u32 g = warp;
      u32 w = sHas[g] & (0xFFFFFFFFu << lane);
      while (!w && ++g < T / 32) w = sHas[g];
      u32 q = 0xFFFFu;
      if (w) {
        q = g * 32 + ((u32)__ffs((int)w) - 1);
        if ((q & 31u) != 31u && ((sHas[q >> 5] >> ((q & 31u) + 1u)) & 1u)) {
          const u32 l1 = (info[q] >> 24) & 0x7Fu, l2 = (info[q + 1] >> 24) & 0x7Fu;
          if (l1 < 8 && l2 >= 8) q++;
        }
      }
      take[tid] = (u16)q;
    }
    __syncthreads();
    if (warp == 0) {
      for (;;) {
        const u32 arrive = anchor > rb ? anchor - rb : 0;
        if (arrive >= T) break;
        const u32 k = take[arrive];
        if (k == 0xFFFFu) break;
        const u32 ee = info[k];
        const u32 mpos = rb + k;
        u32 ml = (ee >> 24) & 0x7Fu;
        const u32 off = ee & 0xFFFFFFu;
        if (ee & kInfoMore) {
          for (;;) {
            const u32 a = mpos + ml + 4 * lane;
            u32 diff = 0;
            if (a + 4 <= bend) {
              diff = gld4(base, fbase + a) ^ gld4(base, fbase + a - off);
            } else {
              for (u32 b = 0; b < 4; b++) {
                if (a + b >= bend || base[fbase + a + b] != base[fbase + a + b - off]) { diff = 0xFFu << (8 * b); break; }
              }
            }
            const u32 bad = __ballot_sync(kFullMask, diff != 0);
            if (!bad) { ml += 128; continue; }
            const u32 first = (u32)__ffs((int)bad) - 1;
            const u32 d = __shfl_sync(kFullMask, diff, first);
            ml += 4 * first + (((u32)__ffs((int)d) - 1) >> 3);
            break;
          }
        }
        const u32 ll = mpos - anchor;
        const u64 rec = emit_sequence(rc, ll, ml, off);
        if (lane == 0) s.seqs[nseq] = rec;
        nseq++;
        litPos += ll;
        anchor = mpos + ml;
      }
      if (lane == 0) sAnchor = anchor;
    }
    __syncthreads();
  }
  // ---- save the tables for the frame's next block
  if (bend < flen) {
    const uint4* z = reinterpret_cast<const uint4*>(smem);
    uint4* gS = reinterpret_cast<uint4*>(s.tabS);
    uint4* gL = reinterpret_cast<uint4*>(s.tabL);
    const u32 nSv = nS >> 2, nLv = nL >> 2;
    for (u32 k = tid; k < nSv; k += T) gS[k] = z[k];
    for (u32 k = tid; k < nLv; k += T) gL[k] = z[nSv + k];
  }
  // ---- literal gather + histograms (whole CTA). Every thread takes a contiguous run of sequences; a block-wide scan of
  // the runs' literal / input byte counts gives each its starting offsets.
  if (tid == 0) {
    sFinal[0] = litPos; sFinal[1] = nseq; sFinal[2] = rc.rep[0]; sFinal[3] = rc.rep[1]; sFinal[4] = rc.rep[2];
    __threadfence_block();
  }
  __syncthreads();
  anchor = sAnchor;
  litPos = sFinal[0];
  nseq = sFinal[1];
  __threadfence();  // the records were written by warp 0 through global memory
  const u32 per = (nseq + T - 1) / T;
  const u32 q0 = tid * per < nseq ? tid * per : nseq, q1 = q0 + per < nseq ? q0 + per : nseq;
  {
    u32 sumL = 0, sumO = 0;
    for (u32 q = q0; q < q1; q++) {
      const u64 rec = s.seqs[q];
      sumL += seq_ll(rec);
      sumO += seq_ll(rec) + seq_ml(rec);
    }
    sScanL[tid] = sumL;
    sScanO[tid] = sumO;
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of T values by one warp, T / 32 values per lane
    u32 accL = 0, accO = 0;
    u32 locL[T / 32], locO[T / 32];
#pragma unroll
    for (u32 k = 0; k < T / 32; k++) {
      locL[k] = accL; locO[k] = accO;
      accL += sScanL[lane * (T / 32) + k];
      accO += sScanO[lane * (T / 32) + k];
    }
    u32 inL = accL, inO = accO;
    for (u32 d = 1; d < 32; d <<= 1) {
      const u32 a = __shfl_up_sync(kFullMask, inL, d), b = __shfl_up_sync(kFullMask, inO, d);
      if (lane >= d) { inL += a; inO += b; }
    }
    const u32 exL = inL - accL, exO = inO - accO;
#pragma unroll
    for (u32 k = 0; k < T / 32; k++) {
      sScanL[lane * (T / 32) + k] = exL + locL[k];
      sScanO[lane * (T / 32) + k] = exO + locO[k];
    }
  }
  __syncthreads();
  {
    u32 to = sScanL[tid], from = bstart + sScanO[tid];
    for (u32 q = q0; q < q1; q++) {
      const u64 rec = s.seqs[q];
      const u32 ll = seq_ll(rec);
      for (u32 k = 0; k < ll; k++) {
        const u8 b = base[fbase + from + k];
        s.lit[to + k] = b;
        atomicAdd(&hist[b], 1u);
      }
      atomicAdd(&cnt[ll_code_fast(ll)], 1u);
      atomicAdd(&cnt[36 + highbit32(seq_off(rec))], 1u);
      atomicAdd(&cnt[68 + ml_code_fast(seq_ml(rec) - 3)], 1u);
      to += ll;
      from += ll + seq_ml(rec);
    }
  }
  const u32 rest = bend - anchor;
  for (u32 q = tid; q < rest; q += T) {
    const u8 b = base[fbase + anchor + q];
    s.lit[litPos + q] = b;
    atomicAdd(&hist[b], 1u);
  }
  __syncthreads();
  for (u32 k = tid; k < 256; k += T) s.hist[k] = hist[k];
  for (u32 k = tid; k < 128; k += T) s.cnt[k] = cnt[k];
  if (tid == 0) {
    gc.blkActive = 1;
    gc.blkPos = bstart;
    gc.blkLen = blen;
    gc.lastBlock = bend >= flen;
    gc.repSave[0] = rep0In; gc.repSave[1] = rep1In; gc.repSave[2] = rep2In;
    gc.rep[0] = sFinal[2]; gc.rep[1] = sFinal[3]; gc.rep[2] = sFinal[4];
    gc.nbSeq = nseq;
    gc.litSize = litPos + rest;
  }
}

__device__ __forceinline__ void bar_sync_n(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int T>
__global__ void __launch_bounds__(T) k_enc_match_cta_pipe(EncJob j, EncView v) {
  extern __shared__ __align__(16) u8 smem[];
  // Producer / consumer form of k_enc_match_cta: warps 1.. (P = T - 32 threads) run the hash / insert / verify / take
  // phases of round r + 1 while warp 0 walks round r; info[], take[] and sHas[] are double-buffered and handed over
  // with named barriers (FULL: producers arrive, the walker waits; EMPTY: the walker arrives, producers wait before
  // they reuse the buffer two rounds later). The round-skip test uses the anchor published after round r - 2, which the
  // EMPTY barrier makes final, so the tables — and the archive — do not depend on timing.
  constexpr u32 P = T - 32;
  constexpr int BAR_PROD = 1, BAR_FULL = 2, BAR_EMPTY = 4;
  __shared__ u32 sAnchorAfter[2];
  __shared__ u32 sSkip[2];
  __shared__ u32 sFinal[8];
  const u32 i = blockIdx.x;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  EncCtx& gc = v.ctx(i);
  const u32 len = gc.srcLen;
  const u32 logS = v.lay.matchLogS, logL = v.lay.matchLogL, mls = v.lay.matchMls;
  const bool dfast = logL != 0;
  const u32 nS = 1u << logS, nL = dfast ? (1u << logL) : 0u;
  u16* tabS = reinterpret_cast<u16*>(smem);
  u16* tabL = tabS + nS;
  u32* info = reinterpret_cast<u32*>(smem + 2ull * (nS + nL));
  u32* hist = info + 2 * T;   // (info is [2][T]) 256 literal counts, then 36 + 32 + 53 code counts
  u32* cnt = hist + 256;
  u16* take = reinterpret_cast<u16*>(cnt + 128);  // [2][T] per position of the round: the match a walk arriving there takes
  __shared__ u32 sHas[2][T / 32];                  // per 32 positions: which of them have a match
  // the walker leaves a round's sequence records here (one shared-memory store each); the producers, which have the
  // slack, copy them out when they next own the buffer (a round of P positions holds at most P / 4 matches)
  constexpr u32 kStage = T / 4;
  __shared__ __align__(8) u64 stageRec[2][kStage];
  __shared__ u32 stageSide[2][kStage];
  __shared__ u32 sRoundSeq[2][2];                  // per buffer: first sequence of its round, one past the last
  const u8* base = j.in;
  const u64 fbase = j.inOff + (u64)i * j.frameSize;
  const EncScratch s = v.frame(i);
  u32* side = reinterpret_cast<u32*>(s.hufOut);  // per sequence: literal source | literal destination << 16
  // ---- clear tables and histograms
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const u32 n16 = (2u * (nS + nL)) >> 4;
    for (u32 k = tid; k < n16; k += T) z[k] = make_uint4(0, 0, 0, 0);
    for (u32 k = tid; k < 256 + 128; k += T) hist[k] = 0;
    if (tid == 0) { sAnchorAfter[0] = sAnchorAfter[1] = 0; sSkip[0] = sSkip[1] = 0; }
  }
  __syncthreads();
  // selector state (meaningful in warp 0, warp-uniform)
  EncCtx rc;
  rc.rep[0] = 1; rc.rep[1] = 4; rc.rep[2] = 8;
  u32 anchor = 0, nseq = 0, litPos = 0;
  const u32 hashEnd = len >= 16 ? len - 8 : 0;  // positions below this can start a match (zstd: ip < iend - 8)
  const u32 rounds = hashEnd ? (hashEnd + P - 1) / P : 0;
  if (warp != 0) {
    const u32 ptid = tid - 32, pwarp = warp - 1;
    for (u32 r = 0, rb = 0; r < rounds; r++, rb += P) {
      const u32 buf = r & 1u;
      if (r >= 2) {
        bar_sync_n(BAR_EMPTY + buf, T);  // the walker is done with this buffer (round r - 2): its records go out
        const u32 q0 = sRoundSeq[buf][0], q1 = sRoundSeq[buf][1];
        for (u32 q = q0 + ptid; q < q1; q += P) { s.seqs[q] = stageRec[buf][q - q0]; side[q] = stageSide[buf][q - q0]; }
      }
    const u32 pos = rb + ptid;
    const u32 curAnchor = sAnchorAfter[buf];  // the anchor after round r - 2 (final: EMPTY barrier above)
    const bool skip = curAnchor >= rb + P;     // the whole round lies inside a match already taken (uniform)
    if (!skip) {
    // ---- A: hash, read the pre-round candidates, then insert. The insert keeps the LOWEST position
    // of this round per cell (16-bit compare-and-swap on the distance from the round base, under
    // which every older entry ranks above the round's own), so the table — and with it the archive —
    // does not depend on thread timing, and later positions of the round see an in-round candidate.
    const bool live = pos < hashEnd;
    u64 x = 0;
    u32 hS = 0, hL = 0, cS = 0, cL = 0;
    if (live) {
      x = gld8(base, fbase + pos);
      hS = mls <= 4 ? ((u32)x * 2654435761u) >> (32 - logS) : (u32)(((x << (64 - 8 * mls)) * 0x9E3779B97F4A7C15ull) >> (64 - logS));
      cS = tabS[hS];
      if (dfast) {
        hL = (u32)((x * 0xCF1BBCDCB7A56463ull) >> (64 - logL));
        cL = tabL[hL];
      }
    }
    bar_sync_n(BAR_PROD, P);
    if (live) {
      table_insert_min(&tabS[hS], pos, rb);
      if (dfast) table_insert_min(&tabL[hL], pos, rb);
    }
    bar_sync_n(BAR_PROD, P);
    u32 e = 0;
    if (live && pos >= curAnchor && pos > 0) {
      u32 c2 = tabS[hS];
      if (c2 < pos && c2 >= rb) cS = c2;
      bool okS = cS < pos, okL = false;
      if (dfast) {
        c2 = tabL[hL];
        if (c2 < pos && c2 >= rb) cL = c2;
        okL = cL < pos;
      }
      // ---- B: verify
      u32 bestLen = 0, bestOff = 0;
      u32 nL8 = 0, nS8 = 0;
      if (okL) nL8 = common8(gld8(base, fbase + cL) ^ x);
      if (okS && (!okL || cS != cL)) nS8 = common8(gld8(base, fbase + cS) ^ x);
      if (nL8 >= 4 && nL8 >= nS8) { bestLen = nL8; bestOff = pos - cL; }
      else if (nS8 >= 4) { bestLen = nS8; bestOff = pos - cS; }
      bool more = false;
      if (bestLen == 8) {
        const u32 cand = pos - bestOff;
        more = true;
        while (bestLen < kLaneLenCap && pos + bestLen + 8 <= len) {
          const u32 c = common8(gld8(base, fbase + pos + bestLen) ^ gld8(base, fbase + cand + bestLen));
          bestLen += c;
          if (c < 8) { more = false; break; }
        }
      }
      if (bestLen >= 4) e = bestOff | (bestLen << 16) | (more ? kInfoMore : 0u);
    }
    info[buf * T + ptid] = e;
    {
      const u32 bal = __ballot_sync(kFullMask, e != 0);
      if (lane == 0) sHas[buf][pwarp] = bal;
    }
    bar_sync_n(BAR_PROD, P);
    // ---- B2 (all threads): which match does a walk take that ARRIVES at position rb + tid? The next position at or
    // after it that has a match, moved one byte on when that match is short (< 8) and its right neighbour's is not
    // (the one-byte lazy step; never across a 32-position group, as in the selection loop this replaces). Doing the
    // search here, in parallel, leaves the serial walk two shared-memory loads per sequence.
    {
      u32 g = pwarp;
      u32 w = sHas[buf][g] & (0xFFFFFFFFu << lane);
      while (!w && ++g < P / 32) w = sHas[buf][g];
      u32 q = 0xFFFFu;
      if (w) {
        q = g * 32 + ((u32)__ffs((int)w) - 1);
        if ((q & 31u) != 31u && ((sHas[buf][q >> 5] >> ((q & 31u) + 1u)) & 1u)) {
          const u32 l1 = (info[buf * T + q] >> 16) & 0xFFu, l2 = (info[buf * T + q + 1] >> 16) & 0xFFu;
          if (l1 < 8 && l2 >= 8) q++;
        }
      }
      take[buf * T + ptid] = (u16)q;
    }
    }
      if (ptid == 0) sSkip[buf] = skip ? 1u : 0u;
      __threadfence_block();
      bar_arrive_n(BAR_FULL + buf, T);
    }
    // the walker's last arrivals (rounds whose buffers nobody reuses) are consumed here so that no barrier is left open
    for (u32 k = rounds >= 2 ? rounds - 2 : 0; k < rounds; k++) {
      const u32 buf = k & 1u;
      bar_sync_n(BAR_EMPTY + buf, T);
      const u32 q0 = sRoundSeq[buf][0], q1 = sRoundSeq[buf][1];
      for (u32 q = q0 + ptid; q < q1; q += P) { s.seqs[q] = stageRec[buf][q - q0]; side[q] = stageSide[buf][q - q0]; }
    }
  } else {
    for (u32 r = 0, rb = 0; r < rounds; r++, rb += P) {
      const u32 buf = r & 1u;
      bar_sync_n(BAR_FULL + buf, T);
      const u32 roundSeq0 = nseq;
      if (!sSkip[buf]) {
      {
        for (;;) {
          const u32 arrive = anchor > rb ? anchor - rb : 0;
          if (arrive >= P) break;
          const u32 k = take[buf * T + arrive];
          if (k == 0xFFFFu) break;
          const u32 ee = info[buf * T + k];
          const u32 gb = rb;
          const u32 mpos = gb + k;
          u32 ml = (ee >> 16) & 0xFFu;
          const u32 off = ee & 0xFFFFu;
          if (ee & kInfoMore) {
            // warp-wide extension: 4 bytes per lane and trip
            for (;;) {
              const u32 a = mpos + ml + 4 * lane;
              u32 diff = 0;  // byte k of diff != 0: byte k differs / is past the end
              if (a + 4 <= len) {
                diff = gld4(base, fbase + a) ^ gld4(base, fbase + a - off);
              } else {
                for (u32 b = 0; b < 4; b++) {
                  if (a + b >= len || base[fbase + a + b] != base[fbase + a + b - off]) { diff = 0xFFu << (8 * b); break; }
                }
              }
              const u32 bad = __ballot_sync(kFullMask, diff != 0);
              if (!bad) { ml += 128; continue; }
              const u32 first = (u32)__ffs((int)bad) - 1;
              const u32 d = __shfl_sync(kFullMask, diff, first);
              ml += 4 * first + (((u32)__ffs((int)d) - 1) >> 3);
              break;
            }
          }
          // ---- emit
          const u32 ll = mpos - anchor;
          const u64 rec = emit_sequence(rc, ll, ml, off);
          if (lane == 0) {
            stageRec[buf][nseq - roundSeq0] = rec;
            stageSide[buf][nseq - roundSeq0] = anchor | (litPos << 16);
          }
          nseq++;
          litPos += ll;
          anchor = mpos + ml;
        }
      }
      }
      if (lane == 0) { sAnchorAfter[buf] = anchor; sRoundSeq[buf][0] = roundSeq0; sRoundSeq[buf][1] = nseq; }
      __threadfence_block();
      bar_arrive_n(BAR_EMPTY + buf, T);
    }
    if (lane == 0) sAnchorAfter[0] = anchor;
  }
  // ---- literal gather + histograms (whole CTA), results
  if (tid == 0) {
    sFinal[0] = litPos; sFinal[1] = nseq; sFinal[2] = rc.rep[0]; sFinal[3] = rc.rep[1]; sFinal[4] = rc.rep[2];
    __threadfence_block();
  }
  __syncthreads();
  anchor = sAnchorAfter[0];
  litPos = sFinal[0];
  nseq = sFinal[1];
  __threadfence();  // the records / side entries were written by warp 0 through global memory
  for (u32 q = tid; q < nseq; q += T) {
    const u64 rec = s.seqs[q];
    const u32 sd = side[q];
    const u32 ll = seq_ll(rec), from = sd & 0xFFFFu, to = sd >> 16;
    for (u32 k = 0; k < ll; k++) {
      const u8 b = base[fbase + from + k];
      s.lit[to + k] = b;
      atomicAdd(&hist[b], 1u);
    }
    atomicAdd(&cnt[ll_code_fast(ll)], 1u);
    atomicAdd(&cnt[36 + highbit32(seq_off(rec))], 1u);
    atomicAdd(&cnt[68 + ml_code_fast(seq_ml(rec) - 3)], 1u);
  }
  const u32 rest = len - anchor;
  for (u32 q = tid; q < rest; q += T) {
    const u8 b = base[fbase + anchor + q];
    s.lit[litPos + q] = b;
    atomicAdd(&hist[b], 1u);
  }
  __syncthreads();
  for (u32 k = tid; k < 256; k += T) s.hist[k] = hist[k];
  for (u32 k = tid; k < 128; k += T) s.cnt[k] = cnt[k];
  if (tid == 0) {
    gc.blkActive = 1;
    gc.blkPos = 0;
    gc.blkLen = len;
    gc.lastBlock = 1;
    gc.repSave[0] = 1; gc.repSave[1] = 4; gc.repSave[2] = 8;
    gc.rep[0] = sFinal[2]; gc.rep[1] = sFinal[3]; gc.rep[2] = sFinal[4];
    gc.nbSeq = nseq;
    gc.litSize = litPos + rest;
  }
}

__global__ void k_enc_literals(EncJob j, EncView v) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.nFrames) return;
  EncCtx& c = v.ctx(i);
  if (!c.blkActive) return;
  EncCtx local = c;
  enc_literals(j.in, j.inOff + (u64)i * j.frameSize, local, v.frame(i));
  c.litSize = local.litSize;
}

__global__ void k_enc_plan(EncJob j, EncView v) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.nFrames) return;
  EncCtx c = v.ctx(i);
  if (!c.blkActive) return;
  enc_plan(c, v.frame(i));
  v.ctx(i) = c;
}

__global__ void k_enc_huf(EncJob j, EncView v) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  u32 i = t >> 2, st = t & 3;
  if (i >= j.nFrames) return;
  EncCtx& c = v.ctx(i);
  if (!c.blkActive || c.litMode != 2 || st >= c.nStreams) return;
  EncCtx local = c;
  enc_huf(local, v.frame(i), st);
  c.hufStreamSize[st] = local.hufStreamSize[st];
}

__global__ void k_enc_seq(EncJob j, EncView v) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.nFrames) return;
  EncCtx& c = v.ctx(i);
  if (!c.blkActive) return;
  EncCtx local = c;
  enc_seq(local, v.frame(i));
  c.seqStreamSize = local.seqStreamSize;
}

// One thread per frame writes the block (serial byte copies of ~1/3 of the frame; the data is L2 hot).
__global__ void k_enc_assemble(EncJob j, EncView v) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.nFrames) return;
  EncCtx c = v.ctx(i);
  if (!c.blkActive) return;
  EncParams p = enc_params(j.level, c.srcLen, j.checksum != 0);
  enc_assemble(j.in, j.inOff + (u64)i * j.frameSize, p, c, v.frame(i), v.out(i));
  v.ctx(i) = c;
}

// ------------------------------------------------------------------------------------------
// FSE sequence bitstream, one warp per frame. The three state chains (LL, OF, ML) are the only
// serial part: lanes 0..2 run them over 32 sequences at a time with the encode tables in shared
// memory and leave (bits, nbBits) per sequence; then all 32 lanes lay their sequence's bit string
// (state bits + extra bits, <= 89 bits) at its prefix-summed bit offset in a shared-memory staging
// window, and whole words leave with coalesced stores. Same bit order as enc_seq (enc_core.cuh),
// i.e. ZSTD_encodeSequences (zstd_compress_sequences.c:286-359).
constexpr u32 kSeqEncWarps = 8;
struct SeqEncSmem {
  FseSymTT tt[128];      // LL 0..35 | OF 36..67 | ML 68..120
  u16 states[1280];      // LL 0..511 | OF 512..767 | ML 768..1279
  u32 chain[3][32];      // bits | nbBits << 16 per sequence of the chunk
  u32 stage[128];        // bit window of the chunk
  u8 codes[3][32];
};

__device__ __forceinline__ void bits_append(u64& lo, u64& hi, u32& n, u32 val, u32 nb) {
  if (n < 64) {
    lo |= (u64)val << n;
    if (n + nb > 64) hi |= (u64)val >> (64 - n);
  } else {
    hi |= (u64)val << (n - 64);
  }
  n += nb;
}

__global__ void __launch_bounds__(kSeqEncWarps * 32) k_enc_seq_warp(EncJob j, EncView v) {
  __shared__ __align__(16) SeqEncSmem sm[kSeqEncWarps];
  const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const u32 i = blockIdx.x * kSeqEncWarps + w;
  if (i >= j.nFrames) return;
  EncCtx& c = v.ctx(i);
  if (!c.blkActive) return;
  const u32 nbSeq = c.nbSeq;
  if (!nbSeq) {
    if (lane == 0) c.seqStreamSize = 0;
    return;
  }
  SeqEncSmem& W = sm[w];
  const EncScratch s = v.frame(i);
  for (u32 k = lane; k < 121; k += 32) W.tt[k] = s.tt[k];
  {
    const u32* src = reinterpret_cast<const u32*>(s.states);
    u32* dst = reinterpret_cast<u32*>(W.states);
    for (u32 k = lane; k < 640; k += 32) dst[k] = src[k];
  }
  const u32 logLL = c.seqLog[0], logOF = c.seqLog[1], logML = c.seqLog[2];
  const u32 ttBase = lane == 0 ? 0u : (lane == 1 ? 36u : 68u);
  const u32 stBase = lane == 0 ? 0u : (lane == 1 ? 512u : 768u);
  const u32 cl = lane < 3 ? lane : 0u;
  u32 state = 0;
  u32 carry = 0, carryBits = 0, outWords = 0;
  u32* out32 = reinterpret_cast<u32*>(s.seqOut);
  const u32 capWords = s.seqOutCap >> 2;
  __syncwarp();
  for (u32 done = 0; done < nbSeq; done += 32) {
    const u32 count = nbSeq - done < 32 ? nbSeq - done : 32;
    const bool valid = lane < count;
    // ---- codes of this chunk (sequence nbSeq-1-done-lane: the stream is written last to first)
    u32 ll = 0, mlb = 0, ov = 1, lc = 0, mc = 0, oc = 0;
    if (valid) {
      const u64 q = s.seqs[nbSeq - 1 - done - lane];
      ll = seq_ll(q); mlb = seq_ml(q) - 3; ov = seq_off(q);
      lc = ll_code_fast(ll); mc = ml_code_fast(mlb); oc = highbit32(ov);
      W.codes[0][lane] = (u8)lc; W.codes[1][lane] = (u8)oc; W.codes[2][lane] = (u8)mc;
    }
    for (u32 k = lane; k < 128; k += 32) W.stage[k] = 0;
    __syncwarp();
    // ---- the three state chains
    if (lane < 3) {
      u32 t = 0;
      if (done == 0) {  // the last sequence's symbols ride in the initial states
        const FseSymTT e = W.tt[ttBase + W.codes[cl][0]];
        const u32 nb = (e.deltaNbBits + (1u << 15)) >> 16;
        const u32 value = (nb << 16) - e.deltaNbBits;
        state = W.states[stBase + (u32)((i32)(value >> nb) + e.deltaFindState)];
        W.chain[cl][0] = 0;
        t = 1;
      }
#pragma unroll 4
      for (; t < count; t++) {
        const FseSymTT e = W.tt[ttBase + W.codes[cl][t]];
        const u32 nb = (state + e.deltaNbBits) >> 16;
        W.chain[cl][t] = (state & ((1u << nb) - 1u)) | (nb << 16);
        state = W.states[stBase + (u32)((i32)(state >> nb) + e.deltaFindState)];
      }
    }
    __syncwarp();
    // ---- this lane's bit string: OF, ML, LL state bits, then LL, ML, OF extra bits
    u64 lo = 0, hi = 0;
    u32 n = 0;
    if (valid) {
      const u32 cLL = W.chain[0][lane], cOF = W.chain[1][lane], cML = W.chain[2][lane];
      bits_append(lo, hi, n, cOF & 0xFFFFu, cOF >> 16);
      bits_append(lo, hi, n, cML & 0xFFFFu, cML >> 16);
      bits_append(lo, hi, n, cLL & 0xFFFFu, cLL >> 16);
      bits_append(lo, hi, n, ll - kLLBase[lc], kLLBits[lc]);
      bits_append(lo, hi, n, mlb + 3 - kMLBase[mc], kMLBits[mc]);
      bits_append(lo, hi, n, ov - (1u << oc), oc);
    }
    u32 incl = n;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) {
      const u32 up = __shfl_up_sync(kFullMask, incl, d);
      if (lane >= d) incl += up;
    }
    const u32 total = __shfl_sync(kFullMask, incl, 31);
    if (n) {
      const u32 pos = carryBits + incl - n;
      const u32 word = pos >> 5, sh = pos & 31u;
      const u32 v0 = (u32)lo, v1 = (u32)(lo >> 32), v2 = (u32)hi;
      const u32 o0 = v0 << sh, o1 = __funnelshift_l(v0, v1, sh), o2 = __funnelshift_l(v1, v2, sh), o3 = __funnelshift_l(v2, 0u, sh);
      if (o0) atomicOr(&W.stage[word], o0);
      if (o1) atomicOr(&W.stage[word + 1], o1);
      if (o2) atomicOr(&W.stage[word + 2], o2);
      if (o3) atomicOr(&W.stage[word + 3], o3);
    }
    if (lane == 0 && carryBits) atomicOr(&W.stage[0], carry);
    __syncwarp();
    const u32 totalBits = carryBits + total;
    const u32 full = totalBits >> 5;
    for (u32 k = lane; k < full; k += 32)
      if (outWords + k < capWords) out32[outWords + k] = W.stage[k];
    carry = W.stage[full];
    carryBits = totalBits & 31u;
    outWords += full;
    __syncwarp();
  }
  // ---- final states (ML, OF, LL), end mark
  const u32 stLL = __shfl_sync(kFullMask, state, 0), stOF = __shfl_sync(kFullMask, state, 1), stML = __shfl_sync(kFullMask, state, 2);
  if (lane == 0) {
    u64 acc = carryBits ? (u64)carry : 0ull;
    u32 nb = carryBits;
    acc |= (u64)(stML & ((1u << logML) - 1u)) << nb; nb += logML;
    acc |= (u64)(stOF & ((1u << logOF) - 1u)) << nb; nb += logOF;
    acc |= (u64)(stLL & ((1u << logLL) - 1u)) << nb; nb += logLL;
    acc |= 1ull << nb; nb += 1;
    const u32 bytes = (nb + 7) >> 3;
    u32 pos = outWords * 4;
    for (u32 k = 0; k < bytes; k++, pos++)
      if (pos < s.seqOutCap) s.seqOut[pos] = (u8)(acc >> (8 * k));
    c.seqStreamSize = pos;
  }
}

// ------------------------------------------------------------------------------------------
// Block assembly, one warp per frame: the decisions of enc_assemble (enc_core.cuh) taken
// warp-uniformly, header bytes by lane 0, every section moved by the whole warp.
__device__ __forceinline__ void warp_move(u8* dst, const u8* src, u32 n, u32 lane) {
  // 32-bit stores to the aligned body of dst, source words funnel-shifted into place
  u32 head = (4u - (u32)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;
  if (head > n) head = n;
  if (lane < head) dst[lane] = src[lane];
  dst += head; src += head; n -= head;
  const u32 words = n >> 2;
  const u32 sh = (u32)(reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
  const u32* sw = reinterpret_cast<const u32*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
  u32* dw = reinterpret_cast<u32*>(dst);
  for (u32 k = lane; k < words; k += 32) {
    const u32 a = sw[k], b = sh ? sw[k + 1] : 0u;
    dw[k] = __funnelshift_r(a, b, sh);
  }
  const u32 done = words << 2, tail = n & 3u;
  if (lane < tail) dst[done + lane] = src[done + lane];
}

__global__ void __launch_bounds__(256) k_enc_assemble_warp(EncJob j, EncView v) {
  const u32 i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= j.nFrames) return;
  EncCtx& gc = v.ctx(i);
  if (!gc.blkActive) return;
  const EncCtx c = gc;
  const EncParams p = enc_params(j.level, c.srcLen, j.checksum != 0);
  const EncScratch s = v.frame(i);
  const u8* in = j.in + j.inOff + (u64)i * j.frameSize;
  u8* out = v.out(i);
  u32 op = c.outPos;
#define ZRA_PUT(b) do { if (lane == 0) out[op] = (u8)(b); op++; } while (0)
  if (c.blkPos == 0) {
    op = 0;
    ZRA_PUT(0x28); ZRA_PUT(0xB5); ZRA_PUT(0x2F); ZRA_PUT(0xFD);
    ZRA_PUT(p.checksum << 2);
    ZRA_PUT((c.windowLog - 10) << 3);
  }
  const u32 n = c.litSize;
  u32 litMode = c.litMode;
  u32 litSection = 0, cLit = 0;
  if (litMode == 2) {
    cLit = c.hufHeaderSize + (c.nStreams == 4 ? 6 : 0);
    for (u32 k = 0; k < c.nStreams; k++) cLit += c.hufStreamSize[k];
    const u32 lh = 3 + (n >= 1024) + (n >= 16384);
    litSection = lh + cLit;
    bool spilled = false;
    for (u32 k = 0; k < c.nStreams; k++) spilled |= c.hufStreamSize[k] > s.hufStride;
    if (spilled || litSection + ((n >> 6) + 2) >= n + 3 || cLit >= (1u << 18) ||
        (c.nStreams == 4 && (c.hufStreamSize[0] > 65535 || c.hufStreamSize[1] > 65535 || c.hufStreamSize[2] > 65535)))
      litMode = 0;
  }
  if (litMode == 0) litSection = (n < 32 ? 1 : (n < 4096 ? 2 : 3)) + n;
  else if (litMode == 1) litSection = (n < 32 ? 1 : (n < 4096 ? 2 : 3)) + 1;
  const u32 seqSection = c.seqHeaderSize + c.seqStreamSize;
  const u32 cSize = litSection + seqSection;
  const u32 minGain = (c.blkLen >> 6) + 2;
  bool raw = c.blkLen < 7 || cSize + minGain >= c.blkLen || cSize >= kBlockSizeMax || c.seqStreamSize > s.seqOutCap;
  // RLE block, the reference's rule (zstd_compress.c:2453-2464): never the first block of a frame, only a block whose
  // sequence store says "maybe" (fewer than 4 sequences and fewer than 10 literals), and then every byte must be equal
  bool rleBlock = false;
  if (c.blkPos != 0 && c.nbSeq < 4 && c.litSize < 10 && c.blkLen > 0) {
    const u8 first = in[c.blkPos];
    bool same = true;
    for (u32 k = lane; k < c.blkLen; k += 32) same &= in[c.blkPos + k] == first;
    rleBlock = __all_sync(0xFFFFFFFFu, same);
  }
  if (rleBlock) raw = false;
  const u32 bsize = (raw || rleBlock) ? c.blkLen : cSize;
  const u32 bh = (c.lastBlock ? 1u : 0u) | ((rleBlock ? 1u : (raw ? 0u : 2u)) << 1) | (bsize << 3);
  ZRA_PUT(bh); ZRA_PUT(bh >> 8); ZRA_PUT(bh >> 16);
  if (rleBlock) {
    ZRA_PUT(in[c.blkPos]);
  } else if (raw) {
    warp_move(out + op, in + c.blkPos, c.blkLen, lane);
    op += c.blkLen;
  } else {
    if (litMode == 2) {
      const u32 single = c.nStreams == 1;
      if (n < 1024) {
        const u32 x = 2u | ((single ? 0u : 1u) << 2) | (n << 4) | (cLit << 14);
        ZRA_PUT(x); ZRA_PUT(x >> 8); ZRA_PUT(x >> 16);
      } else if (n < 16384) {
        const u32 x = 2u | (2u << 2) | (n << 4) | (cLit << 18);
        ZRA_PUT(x); ZRA_PUT(x >> 8); ZRA_PUT(x >> 16); ZRA_PUT(x >> 24);
      } else {
        const u64 x = 2u | (3u << 2) | ((u64)n << 4) | ((u64)cLit << 22);
        for (u32 k = 0; k < 5; k++) ZRA_PUT(x >> (8 * k));
      }
      warp_move(out + op, s.hdr, c.hufHeaderSize, lane);
      op += c.hufHeaderSize;
      if (c.nStreams == 4)
        for (u32 k = 0; k < 3; k++) { ZRA_PUT(c.hufStreamSize[k]); ZRA_PUT(c.hufStreamSize[k] >> 8); }
      for (u32 st = 0; st < c.nStreams; st++) {
        warp_move(out + op, s.hufOut + (u64)st * s.hufStride, c.hufStreamSize[st], lane);
        op += c.hufStreamSize[st];
      }
    } else {
      const u32 type = litMode;  // 0 raw, 1 rle
      if (n < 32) ZRA_PUT(type | (n << 3));
      else if (n < 4096) { const u32 x = type | (1u << 2) | (n << 4); ZRA_PUT(x); ZRA_PUT(x >> 8); }
      else { const u32 x = type | (3u << 2) | (n << 4); ZRA_PUT(x); ZRA_PUT(x >> 8); ZRA_PUT(x >> 16); }
      if (type == 1) ZRA_PUT(s.lit[0]);
      else { warp_move(out + op, s.lit, n, lane); op += n; }
    }
    warp_move(out + op, s.hdr + 256, c.seqHeaderSize, lane);
    op += c.seqHeaderSize;
    warp_move(out + op, s.seqOut, c.seqStreamSize, lane);
    op += c.seqStreamSize;
  }
#undef ZRA_PUT
  if (lane == 0) {
    gc.outPos = op;
    gc.litMode = litMode;
    // a raw or RLE block carries no sequences: the decoder's repeat offsets stay what they were (the reference confirms
    // the block's repcodes only when cSize > 1, zstd_compress.c:2474-2477)
    if (raw || rleBlock) { gc.rep[0] = c.repSave[0]; gc.rep[1] = c.repSave[1]; gc.rep[2] = c.repSave[2]; }
  }
}

// XXH64 of the frame's input by four lanes (one accumulator each), then the final size.
__global__ void k_enc_finish(EncJob j, EncView v, u32* __restrict__ sizes) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  u32 i = t >> 2, q = t & 3;
  bool live = i < j.nFrames;
  u64 acc = xxh_init_acc(q);
  u32 len = 0;
  const u8* p = nullptr;
  if (live && j.checksum) {
    len = v.ctx(i).srcLen;
    p = j.in + j.inOff + (u64)i * j.frameSize;
    u32 stripes = len >> 5;
    if (((uintptr_t)p & 7) == 0) {
      const u64* w = reinterpret_cast<const u64*>(p) + q;
      for (u32 k = 0; k < stripes; k++) acc = xxh_round(acc, __ldg(w + 4 * (u64)k));
    } else {
      const u8* b = p + 8 * q;
      for (u32 k = 0; k < stripes; k++) acc = xxh_round(acc, ld64(b + 32 * (u64)k));
    }
  }
  u32 lane = threadIdx.x & 31, q0 = lane & ~3u;
  u64 v1 = __shfl_sync(kFullMask, acc, q0), v2 = __shfl_sync(kFullMask, acc, q0 + 1), v3 = __shfl_sync(kFullMask, acc, q0 + 2),
      v4 = __shfl_sync(kFullMask, acc, q0 + 3);
  if (!live || q != 0) return;
  EncCtx& c = v.ctx(i);
  u32 op = c.outPos;
  if (j.checksum) {
    u64 h;
    if (len >= 32) {
      h = xxh_rotl(v1, 1) + xxh_rotl(v2, 7) + xxh_rotl(v3, 12) + xxh_rotl(v4, 18);
      h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
    } else {
      h = kXP5;
    }
    u32 x = (u32)xxh_finish(h, len, p + (len & ~31u), len & 31u);
    u8* o = v.out(i);
    o[op++] = (u8)x; o[op++] = (u8)(x >> 8); o[op++] = (u8)(x >> 16); o[op++] = (u8)(x >> 24);
    c.outPos = op;
  }
  sizes[i] = op;
}

// ------------------------------------------------------------------------------------------
static inline u32 div_up(u64 a, u32 b) { return (u32)((a + b - 1) / b); }

size_t encode_scratch_bytes(u32 nFrames, u32 frameSize, u32 lastFrameLen, int level, EncodeLayout* lay) {
  // host-side evaluation of the parameter rule (enc_params is device code)
  auto logs = [&](u32 len, u32* ls, u32* ll) {
    int lv = level == 0 ? 3 : level;
    if (lv < 0) lv = 0;
    if (lv > 4) lv = 4;
    u32 cls = len <= (16u << 10) ? 3 : (len <= (128u << 10) ? 2 : (len <= (256u << 10) ? 1 : 0));
    u32 W, C, H, D;
    if (cls == 3) { W = 14; C = lv == 0 ? 12 : 14; H = lv == 0 ? 13 : 15; D = lv >= 3; }
    else if (cls == 2) { W = 17; u32 c[5] = {12, 12, 13, 15, 16}, h[5] = {12, 13, 15, 16, 17}; C = c[lv]; H = h[lv]; D = lv >= 3; }
    else if (cls == 1) { W = 18; u32 c[5] = {12, 13, 14, 16, 16}, h[5] = {13, 14, 14, 16, 17}; C = c[lv]; H = h[lv]; D = lv >= 2; }
    else { u32 w[5] = {19, 19, 20, 21, 21}, c[5] = {12, 13, 15, 16, 18}, h[5] = {13, 14, 16, 17, 18}; W = w[lv]; C = c[lv]; H = h[lv]; D = lv >= 3; }
    u32 srcLog = 6;
    if (len > 64) { srcLog = 0; while ((1ull << srcLog) < len) srcLog++; }
    if (W > srcLog) W = srcLog;
    if (H > W + 1) H = W + 1;
    if (D && C > W) C = W;
    *ls = D ? C : H;
    *ll = D ? H : 0;
  };
  u32 s1, l1, s2, l2;
  logs(frameSize, &s1, &l1);
  logs(lastFrameLen ? lastFrameLen : frameSize, &s2, &l2);
  u32 ls = s1 > s2 ? s1 : s2, ll = l1 > l2 ? l1 : l2;
  lay->tabSEntries = 1u << ls;
  lay->tabLEntries = ll ? (1u << ll) : 1u;
  // frames of at most 64 KiB take the frame-cooperative matcher (16-bit tables in shared memory)
  {
    // table logs of the shared-memory matcher: the level's own, unless overridden (tuning)
    auto envu = [](const char* k, u32 d) { const char* e = getenv(k); return e ? (u32)strtoul(e, nullptr, 10) : d; };
    int lv = level == 0 ? 3 : level;
    u32 mlsTab[5] = {5, 6, 5, 5, 5};
    // The double-fast levels run with tables a quarter of the reference's size: the matcher inserts EVERY position
    // (the serial reference only the ones it visits), so the smaller tables still find as much, and the number of
    // frames resident per SM (= shared memory per CTA) is what sets the speed. Measured on the bench data, 64 KiB
    // frames, level 3 (profiles/r01g): logs 15/16 -> 4.8 GB/s, archive -0.65 % vs the reference's; 13/14 -> 13.7 GB/s,
    // +0.81 % (text); mixed +0.20 %. The 3 % ratio bound of north_star holds with margin.
    u32 dS = ls, dL = ll;
    if (ll) { dS = ls > 12 ? ls - 2 : (ls > 10 ? 10 : ls); dL = ll > 12 ? ll - 2 : (ll > 10 ? 10 : ll); }
    // fast levels: 2^14 entries whatever the reference's log (13..15). Level 2 / 64 KiB 15 -> 14: 10.7 -> 17.9 GB/s,
    // archive -0.36 % -> +0.67 %; level 1 / 64 KiB 13 -> 14: the reference's repeat-offset probes find 4-5 byte matches
    // this matcher (minimum match 6 there) does not, which cost +2.9 % on text at log 13 — too close to the 3 % bound;
    // at log 14: +0.8 % for 10 % of the speed (profiles/r02l_ratio.jsonl and the run after it)
    else if (ls >= 13) dS = 14;
    lay->matchLogS = envu("ZRA_B200_ENC_LOGS", dS);
    lay->matchLogL = envu("ZRA_B200_ENC_LOGL", dL);
    lay->matchMls = envu("ZRA_B200_ENC_MLS", frameSize <= (16u << 10) ? (lv <= 1 ? 5u : 4u) : mlsTab[lv < 0 ? 0 : (lv > 4 ? 4 : lv)]);
    if (lay->matchLogS > 16) lay->matchLogS = 16;
    if (lay->matchLogL > 16) lay->matchLogL = 16;
    if (lay->matchLogS < 8) lay->matchLogS = 8;
    if (lay->matchLogL && lay->matchLogL < 8) lay->matchLogL = 8;
  }
  lay->matchSmem = 2u * ((1u << lay->matchLogS) + (lay->matchLogL ? (1u << lay->matchLogL) : 0u));
  lay->matchThreads = lay->matchSmem > 100u * 1024u ? 512u : 256u;
  if (getenv("ZRA_B200_ENC_THREADS")) lay->matchThreads = atoi(getenv("ZRA_B200_ENC_THREADS")) >= 512 ? 512u : 256u;
  // producer / consumer form: pays when few frames fit an SM (big tables: level 3 at 64 KiB, 48 KiB -> 4 CTAs per SM:
  // 16.4 -> 20.3 GB/s text, 20.7 -> 24.4 mixed), costs ~10 % when many do (their phases already overlap across CTAs,
  // and the walker warp takes 1/8 of the threads): profiles/r02d
  lay->matchPipe = getenv("ZRA_B200_ENC_PIPE") ? (u32)(atoi(getenv("ZRA_B200_ENC_PIPE")) != 0) : (lay->matchSmem >= 40u * 1024u ? 1u : 0u);
  lay->matchSmem += (lay->matchPipe ? 2u : 1u) * (4u + 2u) * lay->matchThreads + 4u * (256u + 128u);
  lay->ctaMatch = frameSize <= 65536u && lay->matchSmem <= 226u * 1024u && !getenv("ZRA_B200_ENC_SERIAL");
  // frames above 64 KiB (up to 16 MiB): the same matcher with 32-bit tables kept per frame between its blocks
  lay->ctaBig = 0;
  if (!lay->ctaMatch && frameSize > 65536u && frameSize <= (1u << 24) && !getenv("ZRA_B200_ENC_SERIAL")) {
    auto envu = [](const char* k, u32 d) { const char* e = getenv(k); return e ? (u32)strtoul(e, nullptr, 10) : d; };
    // table logs: a big frame has more history to keep than a 64 KiB one, and with EVERY position inserted a small
    // table only remembers the recent past. Measured at 256 KiB frames, 32 MiB of text / mixed data, archive size
    // against the reference's (gpurun_out/r04c, r04d and the sweep after them): double-fast 13/14 -> +3.3 % (L2 text),
    // +2.9 % (L3 mixed) at 12.9 GB/s; 13/15 -> +1.1 %, +2.4 % at 8.7 GB/s; fast 2^14 -> +3.6 % (L1 text) at 14.5 GB/s,
    // 2^15 -> +0.9 % at 9.4 GB/s. The 3 % bound of north_star decides: 13/15 and 2^15 (one CTA per SM).
    u32 bS = ls, bL = ll;
    if (bL) { if (bS > 13) bS = 13; if (bL > 15) bL = 15; if (bL < 15 && ll >= 14) bL = 15; } else bS = 15;
    lay->matchLogS = envu("ZRA_B200_ENC_BIG_LOGS", bS);
    lay->matchLogL = envu("ZRA_B200_ENC_BIG_LOGL", bL);
    lay->matchThreads = 512;
    lay->matchPipe = 0;
    lay->matchSmem = 4u * ((1u << lay->matchLogS) + (lay->matchLogL ? (1u << lay->matchLogL) : 0u)) + 4u * lay->matchThreads + 4u * (256u + 128u) +
                     2u * lay->matchThreads;
    // the HBM slots the tables are parked in between blocks must hold them
    if (lay->tabSEntries < (1u << lay->matchLogS)) lay->tabSEntries = 1u << lay->matchLogS;
    if (lay->matchLogL && lay->tabLEntries < (1u << lay->matchLogL)) lay->tabLEntries = 1u << lay->matchLogL;
    lay->ctaBig = lay->matchSmem <= 200u * 1024u ? 1u : 0u;
  }
  const u32 blk = frameSize < kBlockSizeMax ? frameSize : kBlockSizeMax;
  lay->seqStride = blk / 3 + 2;
  lay->litStride = (blk + 31u) & ~15u;
  lay->hufStride = (((blk + 3) / 4) * 2 + 31u) & ~15u;
  lay->seqOutStride = (blk + 79u) & ~15u;
  u64 bound = (u64)frameSize + (frameSize >> 8) + (frameSize < (128u << 10) ? (((128u << 10) - frameSize) >> 11) : 0);
  lay->outStride = (u32)((bound + 31u) & ~15ull);
  lay->rounds = frameSize ? (frameSize + kBlockSizeMax - 1) / kBlockSizeMax : 1;
  if (!lay->rounds) lay->rounds = 1;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
  const size_t n = nFrames;
  lay->offCtx = take(sizeof(EncCtx) * n);
  lay->offTabS = take(lay->ctaMatch ? 16 : 4ull * lay->tabSEntries * n);
  lay->offTabL = take(lay->ctaMatch ? 16 : 4ull * lay->tabLEntries * n);
  lay->offSeqs = take(8ull * lay->seqStride * n);
  lay->offLit = take((size_t)lay->litStride * n);
  lay->offHist = take(1024ull * n);
  lay->offCodes = take(sizeof(HufCode) * 256ull * n);
  lay->offHuf = take(4ull * lay->hufStride * n);
  lay->offHdr = take(512ull * n);
  lay->offTT = take(sizeof(FseSymTT) * 128ull * n);
  lay->offStates = take(2ull * 1280 * n);
  lay->offSeqOut = take((size_t)lay->seqOutStride * n);
  lay->offCells = take(1024ull * n);
  lay->offCnt = take((lay->ctaMatch || lay->ctaBig) ? 512ull * n : 16);
  lay->offOut = take((size_t)lay->outStride * n);
  lay->offSizes = take(4ull * n);
  lay->offOffsets = take(8ull * (n + 1));
  lay->offBlockSums = take(8ull * (n / 1024 + 2));
  return off;
}

u32 launch_encode_frames(const void* dIn, u64 inOff, u64 inEnd, u32 frameSize, u32 nFrames, int level, bool checksum, void* scratch,
                         const EncodeLayout& lay, cudaStream_t st) {
  if (!nFrames) return 0;
  EncJob j{static_cast<const u8*>(dIn), inOff, inEnd, frameSize, nFrames, level, checksum ? 1u : 0u};
  EncView v{static_cast<u8*>(scratch), lay};
  u8* s = static_cast<u8*>(scratch);
  if (lay.ctaBig) {
    cudaFuncSetAttribute(k_enc_match_cta_big<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.matchSmem);
  } else if (!lay.ctaMatch) {  // hash tables in HBM start empty for every frame
    cudaMemsetAsync(s + lay.offTabS, 0, 4ull * lay.tabSEntries * nFrames, st);
    cudaMemsetAsync(s + lay.offTabL, 0, 4ull * lay.tabLEntries * nFrames, st);
  } else {
    cudaFuncSetAttribute(k_enc_match_cta<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.matchSmem);
    cudaFuncSetAttribute(k_enc_match_cta<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.matchSmem);
    cudaFuncSetAttribute(k_enc_match_cta_pipe<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.matchSmem);
    cudaFuncSetAttribute(k_enc_match_cta_pipe<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.matchSmem);
  }
  const u32 tpb = 64;
  static const bool serialEntropy = getenv("ZRA_B200_ENC_SERIAL_ENTROPY") != nullptr;  // thread-per-frame stages (debugging)
  u32 launches = 0;
  k_enc_begin<<<div_up(nFrames, 128), 128, 0, st>>>(j, v);
  launches++;
  for (u32 r = 0; r < lay.rounds; r++) {
    if (lay.ctaMatch) {
      if (lay.matchPipe) {
        if (lay.matchThreads == 512) k_enc_match_cta_pipe<512><<<nFrames, 512, lay.matchSmem, st>>>(j, v);
        else k_enc_match_cta_pipe<256><<<nFrames, 256, lay.matchSmem, st>>>(j, v);
      } else if (lay.matchThreads == 512) k_enc_match_cta<512><<<nFrames, 512, lay.matchSmem, st>>>(j, v);
      else k_enc_match_cta<256><<<nFrames, 256, lay.matchSmem, st>>>(j, v);
    } else if (lay.ctaBig) {
      k_enc_match_cta_big<512><<<nFrames, 512, lay.matchSmem, st>>>(j, v, r);
    } else {
      k_enc_match<<<div_up(nFrames, tpb), tpb, 0, st>>>(j, v, r);
      k_enc_literals<<<div_up(nFrames, tpb), tpb, 0, st>>>(j, v);
      launches++;
    }
    k_enc_plan<<<div_up(nFrames, tpb), tpb, 0, st>>>(j, v);
    k_enc_huf<<<div_up((u64)nFrames * 4, 128), 128, 0, st>>>(j, v);
    if (serialEntropy) {
      k_enc_seq<<<div_up(nFrames, tpb), tpb, 0, st>>>(j, v);
      k_enc_assemble<<<div_up(nFrames, tpb), tpb, 0, st>>>(j, v);
    } else {
      k_enc_seq_warp<<<div_up(nFrames, kSeqEncWarps), kSeqEncWarps * 32, 0, st>>>(j, v);
      k_enc_assemble_warp<<<div_up((u64)nFrames * 32, 256), 256, 0, st>>>(j, v);
    }
    launches += 5;
  }
  k_enc_finish<<<div_up((u64)nFrames * 4, 128), 128, 0, st>>>(j, v, reinterpret_cast<u32*>(s + lay.offSizes));
  return launches + 1;
}

}  // namespace zrab
