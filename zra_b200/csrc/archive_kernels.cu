// archive_kernels.cu — seek-table construction and header CRC on the device.
//
//   k_scan_block / k_scan_sums / k_scan_apply   exclusive prefix scan of the compressed frame sizes
//   k_write_entries                             40-bit little-endian seek-table entries
//   k_gather_frames                             variable-size frames -> their final, contiguous place
//   k_crc_chunks / k_crc_fold                   CRC-32 of the header's table section
// Replaces the running `outputOffset` bookkeeping and FixedHeader::CalculateHash of the reference
// (source/zra.cpp:96-107, 128-133, 216-231): there the offsets fall out of a serial loop, here
// every frame is compressed independently so its position is a prefix sum.
#include <cuda_runtime.h>

#include "encode_launch.h"
#include "zfmt.cuh"

namespace zrab {

namespace {
constexpr u32 kScanBlock = 1024;

// Per-block exclusive scan (1024 elements per block) + block totals.
__global__ void __launch_bounds__(kScanBlock) k_scan_block(const u32* __restrict__ sizes, u64* __restrict__ offsets,
                                                           u64* __restrict__ blockSums, u32 n) {
  __shared__ u64 warpSums[32];
  u32 i = blockIdx.x * kScanBlock + threadIdx.x;
  u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 v = i < n ? sizes[i] : 0;
  u64 x = v;
#pragma unroll
  for (u32 d = 1; d < 32; d <<= 1) {
    u64 o = __shfl_up_sync(0xFFFFFFFFu, x, d);
    if (lane >= d) x += o;
  }
  if (lane == 31) warpSums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    u64 w = warpSums[lane];
    u64 y = w;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) {
      u64 o = __shfl_up_sync(0xFFFFFFFFu, y, d);
      if (lane >= d) y += o;
    }
    warpSums[lane] = y - w;  // exclusive
    if (lane == 31) blockSums[blockIdx.x] = y;
  }
  __syncthreads();
  if (i < n) offsets[i] = warpSums[warp] + x - v;
}

// Serial exclusive scan of the block totals by one thread (a few thousand values at most).
__global__ void k_scan_sums(u64* blockSums, u32 nBlocks, u64 base, u64* total) {
  u64 run = base == kScanContinue ? *total : base;
  for (u32 b = 0; b < nBlocks; b++) {
    u64 v = blockSums[b];
    blockSums[b] = run;
    run += v;
  }
  *total = run;
}

__global__ void k_scan_apply(u64* __restrict__ offsets, const u64* __restrict__ blockSums, u32 n) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) offsets[i] += blockSums[i / kScanBlock];
  if (i == n) offsets[n] = 0;  // slot n is filled by the caller's total
}

__global__ void k_write_entries(const u64* __restrict__ offsets, u8* __restrict__ table, u64 firstFrame, u32 n,
                                const u64* __restrict__ total) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  u64 v = i < n ? offsets[i] : *total;  // entry n of the batch = running total (sentinel when it is the last batch)
  u8* e = table + 5 * (firstFrame + i);
  e[0] = (u8)v; e[1] = (u8)(v >> 8); e[2] = (u8)(v >> 16); e[3] = (u8)(v >> 24); e[4] = (u8)(v >> 32);
}

// One warp per frame: slot -> final position. The copy is byte granular on the destination side
// (frames land on arbitrary byte offsets), 128 contiguous bytes per warp instruction.
__global__ void __launch_bounds__(256) k_gather_frames(const u8* __restrict__ slots, u32 outStride, const u32* __restrict__ sizes,
                                                       const u64* __restrict__ offsets, u8* __restrict__ dst, u64 dstCap, u32 n) {
  u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const u8* src = slots + (u64)warp * outStride;
  u64 off = offsets[warp];
  u32 len = sizes[warp];
  if (off + len > dstCap) return;  // the caller checks the total against the capacity
  u8* d = dst + off;
  // head up to 4-byte alignment of the destination, then 32-bit words assembled from the slot
  u32 head = (u32)((4 - ((uintptr_t)d & 3)) & 3);
  if (head > len) head = len;
  if (lane < head) d[lane] = src[lane];
  u32 words = (len - head) >> 2;
  for (u32 w = lane; w < words; w += 32) {
    const u8* s = src + head + 4 * w;
    u32 x = (u32)s[0] | ((u32)s[1] << 8) | ((u32)s[2] << 16) | ((u32)s[3] << 24);
    *reinterpret_cast<u32*>(d + head + 4 * w) = x;
  }
  u32 done = head + 4 * words;
  if (done + lane < len) d[done + lane] = src[done + lane];
}

__global__ void k_sizes64(const u32* __restrict__ sizes, u64* __restrict__ out, u32 n) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = sizes[i];
}

// ---------------------------------------------------------------- CRC-32
constexpr u32 kCrcChunk = 4096;

__device__ __forceinline__ u32 crc_table_entry(u32 i) {
  u32 c = i;
#pragma unroll
  for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
  return c;
}

// Each thread: finished CRC (init/xorout all ones) of one 4 KiB chunk, byte-wise through a 256-entry table.
__global__ void __launch_bounds__(256) k_crc_chunks(const u8* __restrict__ data, u64 n, u32* __restrict__ chunkCrc, u32 nChunks) {
  __shared__ u32 table[256];
  table[threadIdx.x] = crc_table_entry(threadIdx.x);
  __syncthreads();
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nChunks) return;
  u64 begin = (u64)i * kCrcChunk;
  u64 end = begin + kCrcChunk < n ? begin + kCrcChunk : n;
  u32 c = 0xFFFFFFFFu;
  for (u64 p = begin; p < end; p++) c = table[(c ^ data[p]) & 0xFFu] ^ (c >> 8);
  chunkCrc[i] = ~c;
}

struct CrcMatrix {
  u32 row[32];
};
__device__ __forceinline__ u32 crc_mat_times(const CrcMatrix& m, u32 vec) {
  u32 sum = 0;
#pragma unroll
  for (int k = 0; k < 32; k++) sum ^= m.row[k] & (0u - ((vec >> k) & 1u));
  return sum;
}

// crc(A || B) = shift(crc(A), |B|) ^ crc(B): fold the chunk CRCs left to right. `full` advances a CRC
// over one whole chunk of zeros, `tail` over the (shorter) last chunk.
__global__ void k_crc_fold(const u32* __restrict__ chunkCrc, u32 nChunks, CrcMatrix full, CrcMatrix tail, u32* __restrict__ out) {
  u32 crc = nChunks ? chunkCrc[0] : 0;
  for (u32 i = 1; i < nChunks; i++) crc = crc_mat_times(i + 1 == nChunks ? tail : full, crc) ^ chunkCrc[i];
  *out = crc;
}

// ---- host-side GF(2) helpers (zlib's crc32_combine construction)
u32 gf2_times(const u32* mat, u32 vec) {
  u32 sum = 0;
  while (vec) {
    if (vec & 1) sum ^= *mat;
    vec >>= 1;
    mat++;
  }
  return sum;
}
void gf2_square(u32* square, const u32* mat) {
  for (int n = 0; n < 32; n++) square[n] = gf2_times(mat, mat[n]);
}
// Operator that advances a CRC over `len` zero bytes.
void crc_zero_operator(u64 len, u32* op) {
  u32 even[32], odd[32];
  odd[0] = 0xEDB88320u;
  u32 row = 1;
  for (int n = 1; n < 32; n++) { odd[n] = row; row <<= 1; }
  gf2_square(even, odd);  // 2 zero bits
  gf2_square(odd, even);  // 4 zero bits
  // identity
  for (int n = 0; n < 32; n++) op[n] = 1u << n;
  auto apply = [&](const u32* m) {
    u32 tmp[32];
    for (int n = 0; n < 32; n++) tmp[n] = gf2_times(m, op[n]);
    for (int n = 0; n < 32; n++) op[n] = tmp[n];
  };
  // odd now = 4 bits; squaring alternately yields 1 byte, 2 bytes, 4 bytes ... operators
  while (len) {
    gf2_square(even, odd);  // even = operator for 2^k bytes (first: 1 byte)
    if (len & 1) apply(even);
    len >>= 1;
    if (!len) break;
    gf2_square(odd, even);
    if (len & 1) apply(odd);
    len >>= 1;
  }
}
}  // namespace

u32 crc32_combine(u32 crcA, u32 crcB, u64 lenB) {
  if (!lenB) return crcA;
  u32 op[32];
  crc_zero_operator(lenB, op);
  return gf2_times(op, crcA) ^ crcB;
}

size_t crc32_workspace_bytes(u64 n) { return 4 * ((n + kCrcChunk - 1) / kCrcChunk) + 64; }

u32 launch_crc32(const u8* dData, u64 n, u32* dCrc, void* workspace, cudaStream_t st) {
  u32 nChunks = (u32)((n + kCrcChunk - 1) / kCrcChunk);
  u32* chunk = static_cast<u32*>(workspace);
  if (!nChunks) {
    cudaMemsetAsync(dCrc, 0, 4, st);
    return 0;
  }
  k_crc_chunks<<<(nChunks + 255) / 256, 256, 0, st>>>(dData, n, chunk, nChunks);
  CrcMatrix full, tail;
  crc_zero_operator(kCrcChunk, full.row);
  u64 last = n - (u64)(nChunks - 1) * kCrcChunk;
  crc_zero_operator(last, tail.row);
  k_crc_fold<<<1, 1, 0, st>>>(chunk, nChunks, full, tail, dCrc);
  return 2;
}

static inline u32 div_up(u64 a, u32 b) { return (u32)((a + b - 1) / b); }

static u32 scan_sizes(void* scratch, const EncodeLayout& lay, u32 nFrames, u64 base, u64* dTotal, cudaStream_t st) {
  u8* s = static_cast<u8*>(scratch);
  const u32* sizes = reinterpret_cast<const u32*>(s + lay.offSizes);
  u64* offsets = reinterpret_cast<u64*>(s + lay.offOffsets);
  u64* sums = reinterpret_cast<u64*>(s + lay.offBlockSums);
  u32 nBlocks = div_up(nFrames, kScanBlock);
  k_scan_block<<<nBlocks, kScanBlock, 0, st>>>(sizes, offsets, sums, nFrames);
  k_scan_sums<<<1, 1, 0, st>>>(sums, nBlocks, base, dTotal);
  k_scan_apply<<<div_up((u64)nFrames + 1, 256), 256, 0, st>>>(offsets, sums, nFrames);
  return 3;
}

u32 launch_scan_gather(void* scratch, const EncodeLayout& lay, u32 nFrames, u64 base, u8* dTable, u64 firstFrame, u8* dFrames,
                       u64 framesCap, u64* dTotal, cudaStream_t st) {
  if (!nFrames) return 0;
  u8* s = static_cast<u8*>(scratch);
  u32 k = scan_sizes(scratch, lay, nFrames, base, dTotal, st);
  const u64* offsets = reinterpret_cast<const u64*>(s + lay.offOffsets);
  k_write_entries<<<div_up((u64)nFrames + 1, 256), 256, 0, st>>>(offsets, dTable, firstFrame, nFrames, dTotal);
  k_gather_frames<<<div_up((u64)nFrames * 32, 256), 256, 0, st>>>(s + lay.offOut, lay.outStride,
                                                                 reinterpret_cast<const u32*>(s + lay.offSizes), offsets, dFrames,
                                                                 framesCap, nFrames);
  return k + 2;
}

u32 launch_scan_pack(void* scratch, const EncodeLayout& lay, u32 nFrames, u8* dOut, u64 outCap, u64* dSizes64, u64* dTotal,
                     cudaStream_t st) {
  if (!nFrames) return 0;
  u8* s = static_cast<u8*>(scratch);
  u32 k = scan_sizes(scratch, lay, nFrames, 0, dTotal, st);
  const u64* offsets = reinterpret_cast<const u64*>(s + lay.offOffsets);
  const u32* sizes = reinterpret_cast<const u32*>(s + lay.offSizes);
  k_gather_frames<<<div_up((u64)nFrames * 32, 256), 256, 0, st>>>(s + lay.offOut, lay.outStride, sizes, offsets, dOut, outCap, nFrames);
  k_sizes64<<<div_up(nFrames, 256), 256, 0, st>>>(sizes, dSizes64, nFrames);
  return k + 2;
}

}  // namespace zrab
