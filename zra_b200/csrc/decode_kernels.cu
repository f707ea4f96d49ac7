// decode_kernels.cu — sm_100a kernels of the frame-parallel zstd decoder and their launcher.
//
// Kernel inventory (one "round" = block r of every frame in the batch; see decode_core.cuh):
//   k_build_descs   1 thread / frame   40-bit seek-table entries -> FrameDesc (ZRA geometry)
//   k_block_setup   1 thread / frame   headers + Huffman/FSE table construction
//   k_huf_decode    1 thread / stream  Huffman literal streams -> literal scratch
//   k_seq_decode    1 thread / frame   FSE sequences -> packed records (+ validation)
//   k_seq_execute   1 warp   / frame   literal/match copies into the output, raw/RLE blocks
//   k_frame_finish  4 threads/ frame   XXH64 checksum + final size checks + error summary
// Replaces the serial loop of ZSTD_decompressMultiFrame that zra::DecompressBuffer /
// DecompressRA / Decompressor / FullDecompressor drive (source/zra.cpp:249,280-293,397-410,435).
#include <cuda_runtime.h>

#include "decode_core.cuh"
#include "decode_launch.h"
#include "xxh64.cuh"

namespace zrab {

constexpr u32 kFull = 0xFFFFFFFFu;
constexpr u32 kLongCopy = 32;  // copies at least this long are done by the whole warp

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

__global__ void k_block_setup(const u8* __restrict__ src, const FrameDesc* __restrict__ descs, FrameCtx* __restrict__ ctxs,
                              FrameTables* __restrict__ tabs, u32 nFrames, u32 firstRound) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nFrames) return;
  FrameCtx c = ctxs[i];
  block_setup(src, descs[i], c, tabs[i], firstRound != 0);
  ctxs[i] = c;
}

__global__ void k_huf_decode(const u8* __restrict__ src, const FrameDesc* __restrict__ descs, FrameCtx* __restrict__ ctxs,
                             const FrameTables* __restrict__ tabs, u8* __restrict__ lit, u32 litStride, u32 nFrames) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  u32 i = t >> 2, s = t & 3;
  if (i >= nFrames) return;
  const FrameCtx& c = ctxs[i];
  if (c.status || c.blkType != BT_COMPRESSED || c.litMode != LIT_HUF || s >= c.nStreams) return;
  u32 e = huf_stream(src, descs[i], c, tabs[i].huf, lit + (u64)i * litStride, s);
  if (e) atomicCAS(&ctxs[i].status, 0u, e);
}

__global__ void k_seq_decode(const u8* __restrict__ src, const FrameDesc* __restrict__ descs, FrameCtx* __restrict__ ctxs,
                             const FrameTables* __restrict__ tabs, u64* __restrict__ seqs, u32 seqStride, u32 nFrames) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nFrames) return;
  FrameCtx c = ctxs[i];
  if (c.blkType != BT_COMPRESSED || c.status) return;
  seq_decode(src, descs[i], c, tabs[i], seqs + (u64)i * seqStride, seqStride);
  // status may have been raised concurrently by k_huf_decode when both run on separate streams;
  // here they are ordered, so a plain write-back of the fields this stage owns is enough
  FrameCtx* g = &ctxs[i];
  if (c.status && !g->status) { g->status = c.status; g->flags = c.flags; g->blkType = c.blkType; }
  g->rep[0] = c.rep[0]; g->rep[1] = c.rep[1]; g->rep[2] = c.rep[2];
  g->blkOut = c.blkOut; g->dstPos = c.dstPos;
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 warp_incl_scan(u32 v, u32 lane) {
#pragma unroll
  for (u32 d = 1; d < 32; d <<= 1) {
    u32 o = __shfl_up_sync(kFull, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

// Whole-warp forward copy, byte granular; no overlap between [dst,dst+n) and [src,src+n).
__device__ __forceinline__ void warp_copy(u8* dst, const u8* src, u32 n, u32 lane) {
  for (u32 i = lane; i < n; i += 32) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) k_seq_execute(const u8* __restrict__ src, u8* dst, const FrameDesc* __restrict__ descs,
                                                     const FrameCtx* __restrict__ ctxs, const u8* __restrict__ lit, u32 litStride,
                                                     const u64* __restrict__ seqs, u32 seqStride, u32 nFrames) {
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
    warp_copy(frame + blkDst, fsrc + c.blkSrc, c.blkSize, lane);
    return;
  }
  if (c.blkType == BT_RLE) {
    u8 v = fsrc[c.blkSrc];
    for (u32 i = lane; i < c.blkSize; i += 32) frame[blkDst + i] = v;
    return;
  }
  // ---- compressed block
  const bool rle = c.litMode == LIT_RLE;
  const u8 rleByte = (u8)c.litSrc;
  const u8* litp = c.litMode == LIT_HUF ? lit + (u64)warp * litStride : fsrc + c.litSrc;
  const u64* sq = seqs + (u64)warp * seqStride;
  const u32 nbSeq = c.nbSeq;
  u32 pos = blkDst;  // frame-relative output cursor
  u32 litPos = 0;
  for (u32 base = 0; base < nbSeq; base += 32) {
    u64 s = (base + lane < nbSeq) ? sq[base + lane] : 0ull;
    u32 ll = seq_ll(s), ml = seq_ml(s), off = seq_off(s);
    u32 sumLL = warp_incl_scan(ll, lane);
    u32 sumOut = warp_incl_scan(ll + ml, lane);
    u32 myLit = litPos + sumLL - ll;
    u32 myDst = pos + sumOut - (ll + ml);
    // literals: long runs by the whole warp, short ones one lane each
    u32 longLit = __ballot_sync(kFull, ll >= kLongCopy);
    while (longLit) {
      int who = __ffs(longLit) - 1;
      longLit &= longLit - 1;
      u32 L = __shfl_sync(kFull, ll, who), from = __shfl_sync(kFull, myLit, who), to = __shfl_sync(kFull, myDst, who);
      if (rle) { for (u32 i = lane; i < L; i += 32) frame[to + i] = rleByte; }
      else warp_copy(frame + to, litp + from, L, lane);
    }
    if (ll < kLongCopy) {
      if (rle) { for (u32 i = 0; i < ll; i++) frame[myDst + i] = rleByte; }
      else { for (u32 i = 0; i < ll; i++) frame[myDst + i] = litp[myLit + i]; }
    }
    __syncwarp();
    // matches: multi-round resolution. Everything below the first pending match is final, so
    // that match can always run; later matches run in the same round when their source lies
    // entirely below it.
    u32 mpos = myDst + ll;
    u32 msrc = mpos - off;
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
          warp_copy(frame + hwm, frame + fs, fml, lane);
        } else {
          // overlapping match: the period [hwm-fo, hwm) is final, every byte is a lookup into it
          for (u32 i = lane; i < fml; i += 32) frame[hwm + i] = frame[fs + (i % fo)];
        }
        if ((int)lane == first) pending = false;
      } else {
        bool ready = pending && ml < kLongCopy && ((int)lane == first || msrc + ml <= hwm);
        if (ready) {
          for (u32 i = 0; i < ml; i++) frame[mpos + i] = frame[msrc + i];
          pending = false;
        }
      }
      __syncwarp();
    }
    pos += __shfl_sync(kFull, sumOut, 31);
    litPos += __shfl_sync(kFull, sumLL, 31);
  }
  // trailing literals
  u32 rest = c.litSize - litPos;
  if (rle) { for (u32 i = lane; i < rest; i += 32) frame[pos + i] = rleByte; }
  else warp_copy(frame + pos, litp + litPos, rest, lane);
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
      const u64* w = reinterpret_cast<const u64*>(p) + q;
      for (u32 k = 0; k < stripes; k++) acc = xxh_round(acc, w[4 * (u64)k]);
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
  return off;
}

void launch_build_descs(const void* archive, u64 tableOff, u64 headerSize, u64 archiveSize, u64 uncompressedSize, u32 frameSize,
                        u32 firstFrame, u32 nFrames, u64 dstBase, void* scratch, const DecodeLayout& lay, cudaStream_t st) {
  if (!nFrames) return;
  u8* s = static_cast<u8*>(scratch);
  k_build_descs<<<div_up(nFrames, 128), 128, 0, st>>>(static_cast<const u8*>(archive), tableOff, headerSize, archiveSize,
                                                      uncompressedSize, frameSize, firstFrame, nFrames, dstBase,
                                                      reinterpret_cast<FrameDesc*>(s + lay.offDescs),
                                                      reinterpret_cast<u32*>(s + lay.offSummary));
}

void launch_summary_reset(void* scratch, const DecodeLayout& lay, cudaStream_t st) {
  // summary = {first failing frame, frames not finished, first bad seek-table entry, rounds}
  static const u32 init[4] = {0xFFFFFFFFu, 0u, 0xFFFFFFFFu, 0u};
  cudaMemcpyAsync(static_cast<u8*>(scratch) + lay.offSummary, init, sizeof(init), cudaMemcpyHostToDevice, st);
}

void launch_decode_rounds(const void* src, void* dst, u32 nFrames, u32 rounds, bool first, void* scratch, const DecodeLayout& lay,
                          cudaStream_t st) {
  if (!nFrames) return;
  u8* s = static_cast<u8*>(scratch);
  auto* descs = reinterpret_cast<FrameDesc*>(s + lay.offDescs);
  auto* ctxs = reinterpret_cast<FrameCtx*>(s + lay.offCtxs);
  auto* tabs = reinterpret_cast<FrameTables*>(s + lay.offTabs);
  u8* lit = s + lay.offLit;
  u64* seqs = reinterpret_cast<u64*>(s + lay.offSeqs);
  const u8* in = static_cast<const u8*>(src);
  for (u32 r = 0; r < rounds; r++) {
    k_block_setup<<<div_up(nFrames, 64), 64, 0, st>>>(in, descs, ctxs, tabs, nFrames, (first && r == 0) ? 1u : 0u);
    k_huf_decode<<<div_up((u64)nFrames * 4, 128), 128, 0, st>>>(in, descs, ctxs, tabs, lit, lay.litStride, nFrames);
    k_seq_decode<<<div_up(nFrames, 64), 64, 0, st>>>(in, descs, ctxs, tabs, seqs, lay.seqStride, nFrames);
    k_seq_execute<<<div_up((u64)nFrames * 32, 256), 256, 0, st>>>(in, static_cast<u8*>(dst), descs, ctxs, lit, lay.litStride, seqs,
                                                                 lay.seqStride, nFrames);
  }
}

void launch_frame_finish(const void* src, const void* dst, u32 nFrames, void* scratch, const DecodeLayout& lay, cudaStream_t st) {
  if (!nFrames) return;
  u8* s = static_cast<u8*>(scratch);
  // "not finished" is recounted by every finish pass
  cudaMemsetAsync(s + lay.offSummary + 4, 0, 4, st);
  k_frame_finish<<<div_up((u64)nFrames * 4, 128), 128, 0, st>>>(static_cast<const u8*>(src), static_cast<const u8*>(dst),
                                                                reinterpret_cast<const FrameDesc*>(s + lay.offDescs),
                                                                reinterpret_cast<FrameCtx*>(s + lay.offCtxs), nFrames,
                                                                reinterpret_cast<u32*>(s + lay.offSummary));
}

u32 frame_status_offset() { return (u32)offsetof(FrameCtx, status); }
u32 frame_ctx_size() { return (u32)sizeof(FrameCtx); }
u32 frame_desc_size() { return (u32)sizeof(FrameDesc); }

}  // namespace zrab
