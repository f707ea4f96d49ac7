/*
 * zra_b200.h — additive device-level C ABI of zra-b200.
 *
 * zra.h keeps the reference's host-pointer contract. The functions here take DEVICE pointers
 * (plain void*, no torch / CUDA types in the signatures; a CUDA stream is passed as void*,
 * NULL = the legacy default stream) so that callers who already hold archives or data in HBM —
 * bench.py, a multi-GPU driver, a storage engine — skip the PCIe staging. Each function names
 * the reference code path it replaces.
 *
 * Alignment rule for every device source pointer: the base address must be 16-byte aligned and
 * the allocation must be readable up to the next multiple of 16 bytes past `size` (the decoder
 * fetches compressed bytes in aligned 16-byte groups; any cudaMalloc / fresh torch allocation
 * satisfies both).
 */
#ifndef ZRA_B200_DEVICE_H
#define ZRA_B200_DEVICE_H

#include "zra.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A GPU context: one CUDA device, its scratch memory and pinned staging. Not thread-safe;
 * use one context per host thread (the zra.h entry points keep one per thread internally).
 * Stands where the reference creates a ZSTD_DCtx / ZSTD_CCtx (source/zra.cpp:43-44). */
typedef struct ZraCudaContext ZraCudaContext;

/* device < 0 selects the calling thread's current CUDA device. */
ZRA_EXPORT ZraStatus ZraCudaCreateContext(ZraCudaContext** context, int device);
ZRA_EXPORT void ZraCudaDestroyContext(ZraCudaContext* context);

/* Text of the last CUDA / decoder failure seen by this context ("" if none). */
ZRA_EXPORT const char* ZraCudaGetLastError(ZraCudaContext* context);

/* Number of kernel launches issued through this context so far (bench.py's gpu_launches). */
ZRA_EXPORT uint64_t ZraCudaGetLaunchCount(ZraCudaContext* context);

/* Per-kernel timing for roofline reports: when enabled, a CUDA event is recorded after every kernel
 * launch and the gaps are attributed per kernel. ZraCudaSetProfiling also clears the totals.
 * ZraCudaGetKernelProfile enumerates index = 1, 2, ... until it returns 0. */
ZRA_EXPORT void ZraCudaSetProfiling(ZraCudaContext* context, int enabled);
ZRA_EXPORT int ZraCudaGetKernelProfile(ZraCudaContext* context, int index, const char** name, double* totalMs,
                                       uint64_t* launches);

/* One independently decodable zstd frame. */
typedef struct ZraCudaFrame {
  uint64_t srcOffset;   /* byte offset of the frame inside the source buffer */
  uint64_t dstOffset;   /* byte offset of its output inside the destination buffer */
  uint32_t srcSize;     /* compressed size */
  uint32_t dstCapacity; /* room for its output */
  uint32_t exact;       /* 1: must regenerate exactly dstCapacity bytes */
  uint32_t reserved;
} ZraCudaFrame;

/* Decodes `count` independent zstd frames (host array `frames`) from dSrc to dDst, both on the
 * device. The GPU form of the reference's per-frame ZSTD_decompressDCtx calls
 * (source/zra.cpp:280,289,293). On failure *failedFrame (may be NULL) is the lowest failing index.
 * frameSizes (host, may be NULL) receives the regenerated size of every frame. */
ZRA_EXPORT ZraStatus ZraCudaDecodeFrames(ZraCudaContext* context, const void* dSrc, size_t srcSize, const ZraCudaFrame* frames,
                                         uint32_t count, void* dDst, uint32_t* frameSizes, uint32_t* failedFrame, void* stream);

/* zra::DecompressBuffer (source/zra.cpp:243-250) with archive and output resident in HBM.
 * outputCapacity must be >= the header's uncompressed size. */
ZRA_EXPORT ZraStatus ZraCudaDecompressBuffer(ZraCudaContext* context, const void* dArchive, size_t archiveSize, void* dOutput,
                                             size_t outputCapacity, void* stream);

/* Frame-range form used for sharding: decodes frames [firstFrame, firstFrame+frameCount)
 * of the archive into dOutput, whose byte 0 corresponds to uncompressed offset firstFrame*frameSize.
 * The multi-GPU equivalent of zra::FullDecompressor::Decompress (source/zra.cpp:428-436). */
ZRA_EXPORT ZraStatus ZraCudaDecompressFrames(ZraCudaContext* context, const void* dArchive, size_t archiveSize,
                                             uint64_t firstFrame, uint64_t frameCount, void* dOutput, size_t outputCapacity,
                                             void* stream);

/* Batched random access (north_star item 4): `count` reads against one device-resident archive in a
 * single call. The reference serves one read per call — zra::DecompressRA (source/zra.cpp:258-296) /
 * zra::Decompressor::Decompress (source/zra.cpp:369-413): frame index = offset / frameSize, seek-table
 * lookup, decode of every touched frame, copy of the slice. Here the requests are mapped to frame
 * indices on the device, frames are DE-DUPLICATED across the batch (each is decoded once), and the
 * requested slices are gathered. All three request arrays are DEVICE pointers:
 *   dOffsets[i]     uncompressed offset of read i
 *   dSizes[i]       its size; NULL = every read is `uniformSize` bytes
 *   dOutOffsets[i]  where its bytes go inside dOutput; NULL = i * uniformSize (needs dSizes == NULL)
 * maxSize bounds every size (ignored when dSizes == NULL). Bounds follow Decompressor::Decompress:
 * offset + size > uncompressedSize is OutOfBoundsAccess and *badRequest (may be NULL) is its index.
 * *uniqueFrames (may be NULL) receives the number of frames actually decoded. */
ZRA_EXPORT ZraStatus ZraCudaDecompressRABatch(ZraCudaContext* context, const void* dArchive, size_t archiveSize,
                                              const uint64_t* dOffsets, const uint32_t* dSizes, const uint64_t* dOutOffsets,
                                              uint32_t uniformSize, uint32_t maxSize, uint64_t count, void* dOutput,
                                              uint64_t* uniqueFrames, uint64_t* badRequest, void* stream);

/* zra::CompressBuffer (source/zra.cpp:194-234) with input and archive resident in HBM: every frame
 * is compressed by the GPU encoder (zstd levels 1-3 semantics; level 0 = 3, higher levels use the
 * strongest implemented parser), the 40-bit seek table is the device prefix scan of the frame sizes
 * and the header CRC-32 is computed on the device. outputCapacity must be at least
 * ZraGetCompressedOutputBufferSize(inputSize, frameSize) + metaSize. Unlike the host entry point this
 * one stores the metadata bytes (it does not reproduce the reference's metadata quirk).
 * The input allocation must be readable up to 8 bytes past inputSize. */
ZRA_EXPORT ZraStatus ZraCudaCompressBuffer(ZraCudaContext* context, const void* dInput, size_t inputSize, void* dOutput,
                                           size_t outputCapacity, size_t* outputSize, int8_t compressionLevel, uint32_t frameSize,
                                           bool checksum, const void* metaBuffer, size_t metaSize, void* stream);

/* ---- sharding across GPUs ---------------------------------------------------------
 * Frames are independent, so GPU g of G owns the contiguous frame range [g*F/G, (g+1)*F/G).
 * Decompression of a shard is ZraCudaDecompressFrames above (no exchange). Compression of a shard is
 * ZraCudaCompressFrames: the shard's frames back to back plus their sizes; the ONE exchange step is
 * the gather of the per-frame sizes (equivalently: the exclusive scan of the per-shard totals gives
 * every shard its base offset), after which ZraShardBuildHeader serialises the archive header. */

/* zra::Compressor::Compress (source/zra.cpp:319-350) for a device-resident shard: dInput[0, inputSize)
 * is cut into frameSize frames (only the last may be short), each becomes one zstd frame, written
 * back to back at dOutput. frameSizes (HOST array, ceil(inputSize / frameSize) entries) receives the
 * compressed size of every frame, *outputSize the total. outputCapacity >= ZSTD_compressBound(frameSize)
 * * frames is always enough. The input must be readable up to 8 bytes past inputSize. */
ZRA_EXPORT ZraStatus ZraCudaCompressFrames(ZraCudaContext* context, const void* dInput, size_t inputSize, uint32_t frameSize,
                                           int8_t compressionLevel, bool checksum, void* dOutput, size_t outputCapacity,
                                           uint64_t* frameSizes, size_t* outputSize, void* stream);

/* Bytes of the header of an archive with `frames` frames and metaSize bytes of metadata. Host only. */
ZRA_EXPORT size_t ZraShardHeaderSize(uint64_t frames, size_t metaSize);

/* Serialises FixedHeader + meta + 40-bit seek table + CRC-32 (source/zra.cpp:111-134, 201-231) from the
 * compressed size of every frame in archive order (the exclusive scan into offsets happens here).
 * Host only, needs no GPU. Fails with CompressedSizeTooLarge when the total reaches 2^40 and with
 * OutputBufferTooSmall when outCapacity < ZraShardHeaderSize(frames, metaSize). */
ZRA_EXPORT ZraStatus ZraShardBuildHeader(uint64_t uncompressedSize, uint32_t frameSize, const void* metaBuffer, size_t metaSize,
                                         const uint64_t* frameSizes, uint64_t frames, void* out, size_t outCapacity);

/* ---- integrity beyond the reference ----------------------------------------------
 * The reference stores a CRC-32 of the header (everything but the hash field itself: fixed header, metadata, seek
 * table; source/zra.cpp:128-133) and never checks it (zra::Header, zra.cpp:141-171). This recomputes it from the
 * first headerSize bytes of an archive in HOST memory: Success when it matches, HeaderInvalid when the magic /
 * version are wrong or the CRC differs, OutOfBoundsAccess when `size` is shorter than the header. Needs no GPU. */
ZRA_EXPORT ZraStatus ZraVerifyHeaderCrc(const void* archive, size_t size);

#ifdef __cplusplus
}
#endif
#endif /* ZRA_B200_DEVICE_H */
