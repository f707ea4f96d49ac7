/*
 * zra.h — C interface of zra-b200, a B200-native implementation of the ZRA
 * ("ZStandard Random Access") archive hot path.
 *
 * Drop-in contract: every type and function below has the name, argument order
 * and meaning of the reference library's C interface (reference:
 * /root/reference/include/zra.h:27-263, implemented at source/zra.cpp:439-625),
 * so a program written against the reference links against libzra_b200.so
 * unchanged. All buffers are HOST pointers; the library stages them to the GPU,
 * runs the CUDA kernels and copies results back. Device-pointer entry points
 * are additive and live in zra_b200.h. There is no CPU fallback: without a
 * usable CUDA device every call that needs one reports ZStdError.
 */
#ifndef ZRA_B200_ZRA_H
#define ZRA_B200_ZRA_H

#if defined(ZRA_EXPORT_HEADER)
#include "zra_export.h"
#elif !defined(ZRA_EXPORT)
#if defined(_WIN32)
#define ZRA_EXPORT __declspec(dllimport)
#else
#define ZRA_EXPORT __attribute__((visibility("default")))
#endif
#endif

#ifdef __cplusplus
#include <cstddef>
#include <cstdint>
extern "C" {
#else
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#endif

/* Outcome classes. Numeric values match zra::StatusCode (zra.hpp) one to one. */
typedef enum ZraStatusCode {
  Success,                /* nothing went wrong */
  ZStdError,              /* the zstd layer (here: the GPU codec) failed; ZraStatus.zstd holds the ZSTD_ErrorCode */
  ZraVersionLow,          /* archive written by a newer format version */
  HeaderInvalid,          /* magic / version check of the archive header failed */
  HeaderIncomplete,       /* streaming compressor: header asked for before the last frame was written */
  OutOfBoundsAccess,      /* requested range lies outside the data */
  OutputBufferTooSmall,   /* caller's output buffer cannot hold the result */
  CompressedSizeTooLarge, /* compressed payload would not fit the 40-bit seek table */
  InputFrameSizeMismatch, /* streaming compressor: non-final chunk not a multiple of the frame size */
} ZraStatusCode;

typedef struct ZraStatus {
  ZraStatusCode zra; /* ZRA-level outcome */
  int zstd;          /* ZSTD_ErrorCode when zra == ZStdError, else 0 */
} ZraStatus;

/* ---- library ---- */

/* Highest archive format version this library reads and writes (1). */
/* reference: include/zra.h:56, source/zra.cpp:449 */
ZRA_EXPORT uint16_t ZraGetVersion();

/* Human-readable text for a status; the pointer refers to static storage. */
/* reference: include/zra.h:62, source/zra.cpp:453 */
ZRA_EXPORT const char* ZraGetErrorString(ZraStatus status);

/* ---- header ---- */
typedef struct ZraHeader ZraHeader;

/* Parses the header through a read callback: readFunction(offset, size, buffer) must fill
 * `buffer` with `size` archive bytes starting at `offset`. */
/* reference: include/zra.h:72, source/zra.cpp:457 */
ZRA_EXPORT ZraStatus ZraCreateHeader(ZraHeader** header, void(readFunction)(size_t offset, size_t size, void* buffer));

/* Parses the header of an archive that is completely in memory. */
/* reference: include/zra.h:80, source/zra.cpp:466 */
ZRA_EXPORT ZraStatus ZraCreateHeader2(ZraHeader** header, void* buffer, size_t size);

/* reference: include/zra.h:85, source/zra.cpp:475 */
ZRA_EXPORT void ZraDeleteHeader(ZraHeader* header);

/* reference: include/zra.h:90, source/zra.cpp:479 */
ZRA_EXPORT size_t ZraGetVersionWithHeader(ZraHeader* header);
/* reference: include/zra.h:95, source/zra.cpp:483 */
ZRA_EXPORT size_t ZraGetHeaderSizeWithHeader(ZraHeader* header);
/* reference: include/zra.h:100, source/zra.cpp:487 */
ZRA_EXPORT size_t ZraGetUncompressedSizeWithHeader(ZraHeader* header);
/* reference: include/zra.h:105, source/zra.cpp:491 */
ZRA_EXPORT size_t ZraGetFrameSizeWithHeader(ZraHeader* header);
/* reference: include/zra.h:110, source/zra.cpp:495 */
ZRA_EXPORT size_t ZraGetMetadataSize(ZraHeader* header);

/* Copies the metadata section into `buffer` (ZraGetMetadataSize bytes). */
/* reference: include/zra.h:115, source/zra.cpp:499 */
ZRA_EXPORT void ZraGetMetadata(ZraHeader* header, void* buffer);

/* ---- whole buffers ---- */

/* Worst-case archive size for `inputSize` bytes cut into `frameSize` frames (no metadata). */
/* reference: include/zra.h:124, source/zra.cpp:504 */
ZRA_EXPORT size_t ZraGetCompressedOutputBufferSize(size_t inputSize, size_t frameSize);

/* Compresses inputBuffer into a complete archive. outputBuffer must hold
 * ZraGetCompressedOutputBufferSize(inputSize, frameSize) bytes; *outputSize receives the archive size. */
/* reference: include/zra.h:138, source/zra.cpp:508 */
ZRA_EXPORT ZraStatus ZraCompressBuffer(void* inputBuffer, size_t inputSize, void* outputBuffer, size_t* outputSize,
                                       int8_t compressionLevel, uint32_t frameSize, bool checksum, void* metaBuffer,
                                       size_t metaSize);

/* Decompresses a complete archive; outputBuffer must hold the header's uncompressed size. */
/* reference: include/zra.h:146, source/zra.cpp:517 */
ZRA_EXPORT ZraStatus ZraDecompressBuffer(void* inputBuffer, size_t inputSize, void* outputBuffer);

/* Decompresses `size` bytes starting at uncompressed position `offset`. */
/* reference: include/zra.h:156, source/zra.cpp:526 */
ZRA_EXPORT ZraStatus ZraDecompressRA(void* inputBuffer, size_t inputSize, void* outputBuffer, size_t offset, size_t size);

/* ---- streaming compressor ---- */
typedef struct ZraCompressor ZraCompressor;

/* `size` is the exact total length of the stream that will be fed in. */
/* reference: include/zra.h:171, source/zra.cpp:535 */
ZRA_EXPORT ZraStatus ZraCreateCompressor(ZraCompressor** compressor, size_t size, int8_t compressionLevel, uint32_t frameSize,
                                         bool checksum, void* metaBuffer, size_t metaSize);
/* reference: include/zra.h:176, source/zra.cpp:544 */
ZRA_EXPORT void ZraDeleteCompressor(ZraCompressor* compressor);

/* Worst-case output of one ZraCompressWithCompressor call fed `inputSize` bytes. */
/* reference: include/zra.h:183, source/zra.cpp:548 */
ZRA_EXPORT size_t ZraGetOutputBufferSizeWithCompressor(ZraCompressor* compressor, size_t inputSize);

/* Compresses the next chunk (a multiple of the frame size unless it is the last one). */
/* reference: include/zra.h:192, source/zra.cpp:552 */
ZRA_EXPORT ZraStatus ZraCompressWithCompressor(ZraCompressor* compressor, void* inputBuffer, size_t inputSize,
                                               void* outputBuffer, size_t* outputSize);

/* reference: include/zra.h:197, source/zra.cpp:562 */
ZRA_EXPORT size_t ZraGetHeaderSizeWithCompressor(ZraCompressor* compressor);

/* Copies the finished header (fixed part, metadata, seek table) into outputBuffer. */
/* reference: include/zra.h:203, source/zra.cpp:566 */
ZRA_EXPORT ZraStatus ZraGetHeaderWithCompressor(ZraCompressor* compressor, void* outputBuffer);

/* ---- streaming random-access decompressor ---- */
typedef struct ZraDecompressor ZraDecompressor;

/* reference: include/zra.h:215, source/zra.cpp:576 */
ZRA_EXPORT ZraStatus ZraCreateDecompressor(ZraDecompressor** decompressor,
                                           void(readFunction)(size_t offset, size_t size, void* buffer), size_t maxCacheSize);
/* reference: include/zra.h:220, source/zra.cpp:585 */
ZRA_EXPORT void ZraDeleteDecompressor(ZraDecompressor* decompressor);

/* Borrowed pointer: valid until the decompressor is deleted; do not ZraDeleteHeader it. */
/* reference: include/zra.h:226, source/zra.cpp:589 */
ZRA_EXPORT ZraHeader* ZraGetHeaderWithDecompressor(ZraDecompressor* decompressor);

/* reference: include/zra.h:234, source/zra.cpp:593 */
ZRA_EXPORT ZraStatus ZraDecompressWithDecompressor(ZraDecompressor* decompressor, size_t offset, size_t size,
                                                   void* outputBuffer);

/* ---- streaming whole-archive decompressor ---- */
typedef struct ZraFullDecompressor ZraFullDecompressor;

/* reference: include/zra.h:244, source/zra.cpp:602 */
ZRA_EXPORT ZraStatus ZraCreateFullDecompressor(ZraFullDecompressor** decompressor,
                                               void(readFunction)(size_t offset, size_t size, void* buffer),
                                               size_t maxCacheSize);
/* reference: include/zra.h:249, source/zra.cpp:611 */
ZRA_EXPORT void ZraDeleteFullDecompressor(ZraFullDecompressor* decompressor);

/* Borrowed pointer, same rules as ZraGetHeaderWithDecompressor. */
/* reference: include/zra.h:255, source/zra.cpp:615 */
ZRA_EXPORT ZraHeader* ZraGetHeaderWithFullDecompressor(ZraFullDecompressor* decompressor);

/* Decompresses as many whole frames as fit outputCapacity; *outputSize == 0 means the end was reached. */
/* reference: include/zra.h:263, source/zra.cpp:619 */
ZRA_EXPORT ZraStatus ZraDecompressWithFullDecompressor(ZraFullDecompressor* decompressor, void* outputBuffer,
                                                       size_t outputCapacity, size_t* outputSize);

#ifdef __cplusplus
}
#endif
#endif /* ZRA_B200_ZRA_H */
