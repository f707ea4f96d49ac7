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
  lay->offTabS = take(4ull * lay->tabSEntries * n);
  lay->offTabL = take(4ull * lay->tabLEntries * n);
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
  // hash tables start empty for every frame
  cudaMemsetAsync(s + lay.offTabS, 0, 4ull * lay.tabSEntries * nFrames, st);
  cudaMemsetAsync(s + lay.offTabL, 0, 4ull * lay.tabLEntries * nFrames, st);
  const u32 tpb = 64;
  u32 launches = 0;
  k_enc_begin<<<div_up(nFrames, 128), 128, 0, st>>>(j, v);
  launches++;
  for (u32 r = 0; r < lay.rounds; r++) {
    k_enc_match<<<div_up(nFrames, tpb), tpb, 0, st>>>(j, v, r);
    k_enc_literals<<<div_up(nFrames, tpb), tpb, 0, st>>>(j, v);
    k_enc_plan<<<div_up(nFrames, tpb), tpb, 0, st>>>(j, v);
    k_enc_huf<<<div_up((u64)nFrames * 4, 128), 128, 0, st>>>(j, v);
    k_enc_seq<<<div_up(nFrames, tpb), tpb, 0, st>>>(j, v);
    k_enc_assemble<<<div_up(nFrames, tpb), tpb, 0, st>>>(j, v);
    launches += 6;
  }
  k_enc_finish<<<div_up((u64)nFrames * 4, 128), 128, 0, st>>>(j, v, reinterpret_cast<u32*>(s + lay.offSizes));
  return launches + 1;
}

}  // namespace zrab
