// ra_kernels.cu — batched random access (north_star item 4).
//
// The reference serves one request at a time: frame index = offset / frameSize, seek-table lookup,
// decode 1..n whole frames, copy the slice (source/zra.cpp:258-296, 369-413). Here a whole batch of
// requests is mapped at once:
//   k_ra_map     1 thread / request   bounds check, frame range, and CLAIM of every touched frame:
//                                     the first toucher of a frame (atomicCAS on slotOf[frame])
//                                     takes the next slot, so every frame is decoded once per batch
//   k_ra_descs   1 thread / slot      seek-table entries -> FrameDesc (output = slot * frameSize)
//   (frame decode: decode_kernels.cu, unchanged)
//   k_ra_gather  1 warp / request     requested slice out of the decoded slots, 16-byte stores
//   k_ra_reset   1 thread / slot      slotOf[] back to "empty" for the next batch
#include <cuda_runtime.h>

#include "decode_core.cuh"
#include "ra_launch.h"

namespace zrab {

constexpr u32 kEmpty = 0xFFFFFFFFu, kClaimed = 0xFFFFFFFEu;

__device__ __forceinline__ u64 ra_size(const RaBatch& b, u64 i) { return b.sizes ? b.sizes[i] : b.uniformSize; }

__global__ void k_ra_map(RaBatch b, u64 first, u32 n, u32* __restrict__ slotOf, u32* __restrict__ uniqueFrames,
                         u32* __restrict__ counters) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  u64 i = first + t;
  u64 off = b.offsets[i], size = ra_size(b, i);
  // zra::Decompressor::Decompress bound (source/zra.cpp:370): offset + size > uncompressedSize is out of bounds
  if (off > b.uncompressedSize || size > b.uncompressedSize - off) {
    atomicMin(&counters[1], t);
    return;
  }
  if (!size) return;
  u64 f0 = off / b.frameSize, f1 = (off + size - 1) / b.frameSize;
  for (u64 f = f0; f <= f1; f++) {
    if (atomicCAS(&slotOf[f], kEmpty, kClaimed) == kEmpty) {
      u32 slot = atomicAdd(&counters[0], 1u);
      uniqueFrames[slot] = (u32)f;
      slotOf[f] = slot;  // read by k_ra_descs / k_ra_gather, i.e. after this kernel has finished
    }
  }
}

__global__ void k_ra_descs(const u8* __restrict__ archive, u64 tableOff, u64 headerSize, u64 archiveSize, u64 uncompressedSize,
                           u32 frameSize, const u32* __restrict__ uniqueFrames, u32 nUnique, FrameDesc* __restrict__ descs) {
  u32 slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nUnique) return;
  u64 f = uniqueFrames[slot];
  const u8* e = archive + tableOff + 5 * f;
  u64 a = (u64)ld32(e) | ((u64)e[4] << 32);
  u64 b = (u64)ld32(e + 5) | ((u64)e[9] << 32);
  FrameDesc d;
  u64 begin = f * frameSize, left = uncompressedSize - begin;
  d.srcOff = headerSize + a;
  d.dstOff = (u64)slot * frameSize;
  d.dstCap = (u32)(left < frameSize ? left : frameSize);
  d.exact = 1;
  d.pad = 0;
  d.srcLen = (b < a || headerSize + b > archiveSize || b - a > 0xFFFFFFFFull) ? 0u : (u32)(b - a);  // 0 decodes to srcSize_wrong
  descs[slot] = d;
}

__global__ void __launch_bounds__(256) k_ra_gather(RaBatch b, u64 first, u32 n, const u32* __restrict__ slotOf,
                                                  const u8* __restrict__ slots, u8* __restrict__ out) {
  u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  u64 i = first + w;
  u64 off = b.offsets[i], size = ra_size(b, i);
  u8* dst = out + (b.outOffsets ? b.outOffsets[i] : i * b.uniformSize);
  while (size) {
    u64 f = off / b.frameSize;
    u32 in = (u32)(off - f * b.frameSize);
    u32 len = (u32)(size < b.frameSize - in ? size : b.frameSize - in);
    const u8* src = slots + (u64)slotOf[f] * b.frameSize + in;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      for (u32 v = lane; v < (len >> 4); v += 32) d4[v] = s4[v];
      for (u32 k = (len & ~15u) + lane; k < len; k += 32) dst[k] = src[k];
    } else {
      for (u32 k = lane; k < len; k += 32) dst[k] = src[k];
    }
    off += len; size -= len; dst += len;
  }
}

__global__ void k_ra_reset(u32* __restrict__ slotOf, const u32* __restrict__ uniqueFrames, u32 nUnique) {
  u32 slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot < nUnique) slotOf[uniqueFrames[slot]] = kEmpty;
}

static inline u32 div_up(u64 a, u32 b) { return (u32)((a + b - 1) / b); }

void launch_ra_map(const RaBatch& b, u64 first, u32 n, u32* slotOf, u32* uniqueFrames, u32* counters, cudaStream_t st) {
  // counters = {unique frames, first out-of-bounds request}
  cudaMemsetAsync(counters, 0, 4, st);
  cudaMemsetAsync(counters + 1, 0xFF, 4, st);
  if (n) k_ra_map<<<div_up(n, 256), 256, 0, st>>>(b, first, n, slotOf, uniqueFrames, counters);
}
void launch_ra_descs(const void* archive, u64 tableOff, u64 headerSize, u64 archiveSize, u64 uncompressedSize, u32 frameSize,
                     const u32* uniqueFrames, u32 nUnique, void* descs, cudaStream_t st) {
  if (nUnique)
    k_ra_descs<<<div_up(nUnique, 256), 256, 0, st>>>(static_cast<const u8*>(archive), tableOff, headerSize, archiveSize,
                                                    uncompressedSize, frameSize, uniqueFrames, nUnique,
                                                    static_cast<FrameDesc*>(descs));
}
void launch_ra_gather(const RaBatch& b, u64 first, u32 n, const u32* slotOf, const void* slots, void* out, cudaStream_t st) {
  if (n) k_ra_gather<<<div_up((u64)n * 32, 256), 256, 0, st>>>(b, first, n, slotOf, static_cast<const u8*>(slots), static_cast<u8*>(out));
}
void launch_ra_reset(u32* slotOf, const u32* uniqueFrames, u32 nUnique, cudaStream_t st) {
  if (nUnique) k_ra_reset<<<div_up(nUnique, 256), 256, 0, st>>>(slotOf, uniqueFrames, nUnique);
}

}  // namespace zrab
