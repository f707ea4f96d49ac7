// encode_launch.h — host-visible launch interface of encode_kernels.cu and archive_kernels.cu (internal).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace zrab {

// Per-batch device scratch of the encoder. Every frame in flight owns one slice of each region.
struct EncodeLayout {
  size_t offCtx, offTabS, offTabL, offSeqs, offLit, offHist, offCodes, offHuf, offHdr, offTT, offStates, offSeqOut, offCells, offCnt,
      offOut, offSizes, offOffsets, offBlockSums;
  uint32_t tabSEntries, tabLEntries;  // per frame
  uint32_t seqStride, litStride, hufStride, seqOutStride, outStride;
  uint32_t rounds;                    // blocks per frame
  uint32_t matchPipe;                          // producer / consumer form of the matcher (experiment)
  uint32_t ctaMatch, matchSmem, matchThreads;  // frame-cooperative matcher (frames <= 64 KiB)
  uint32_t ctaBig;                             // its 32-bit form for frames above 64 KiB (tables parked in HBM between blocks)
  uint32_t matchLogS, matchLogL, matchMls;     // its table logs (16-bit entries) and short-hash width
};

// Scratch needed to encode `nFrames` frames of at most `frameSize` bytes at `level`.
size_t encode_scratch_bytes(uint32_t nFrames, uint32_t frameSize, uint32_t lastFrameLen, int level, EncodeLayout* lay);

// Encodes frames [0, nFrames) of the input range starting at byte `inOff` of dIn (frame i covers
// [inOff + i*frameSize, +min(frameSize, inEnd - ...))). Results: sizes[i] (u32, at offSizes) and the
// frame bytes at offOut + i*outStride. Returns the number of kernels launched.
uint32_t launch_encode_frames(const void* dIn, uint64_t inOff, uint64_t inEnd, uint32_t frameSize, uint32_t nFrames, int level,
                              bool checksum, void* scratch, const EncodeLayout& lay, cudaStream_t st);

// Exclusive scan of the batch's frame sizes (+ `base`), 40-bit seek-table entries at
// dTable + 5*firstFrame, frames gathered to dFrames + offset. *dTotal (device u64) = base + batch total.
// base == kScanContinue: the base is what *dTotal holds when the scan runs (batches queued back to back).
constexpr uint64_t kScanContinue = ~0ull;
uint32_t launch_scan_gather(void* scratch, const EncodeLayout& lay, uint32_t nFrames, uint64_t base, uint8_t* dTable,
                            uint64_t firstFrame, uint8_t* dFrames, uint64_t framesCap, uint64_t* dTotal, cudaStream_t st);

// Streaming variant: frames gathered back to back into dOut (no table), per-frame sizes copied to dSizes64.
uint32_t launch_scan_pack(void* scratch, const EncodeLayout& lay, uint32_t nFrames, uint8_t* dOut, uint64_t outCap,
                          uint64_t* dSizes64, uint64_t* dTotal, cudaStream_t st);

// CRC-32 (zlib polynomial) of dData[0..n) computed on the device; the finished CRC lands in *dCrc.
// `workspace` needs 4 * ceil(n / 4096) + 64 bytes.
uint32_t launch_crc32(const uint8_t* dData, uint64_t n, uint32_t* dCrc, void* workspace, cudaStream_t st);
size_t crc32_workspace_bytes(uint64_t n);

// Host helpers for chaining CRCs (crc(A || B) from crc(A), crc(B), |B|).
uint32_t crc32_combine(uint32_t crcA, uint32_t crcB, uint64_t lenB);

}  // namespace zrab
