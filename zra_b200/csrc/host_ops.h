// host_ops.h — host-buffer operations behind the zra:: API (internal): stage, launch, copy back.
#pragma once
#include <functional>
#include <cstddef>
#include <cstdint>

#include "gpu_context.h"

namespace zrab {

struct OpStatus {
  int zra{0};   // zra::StatusCode value
  int zstd{0};  // ZSTD_ErrorCode when zra == 1
  bool cuda{false};
};

// Whole archive (host) -> out (host, >= info.uncompressedSize bytes).
OpStatus host_decompress_archive(GpuContext* g, const uint8_t* archive, size_t n, const ArchiveInfo& info, uint8_t* out);

// Bytes [offset, offset+size) of the original data of an in-memory archive.
OpStatus host_decompress_range(GpuContext* g, const uint8_t* archive, size_t n, const ArchiveInfo& info, uint64_t offset,
                               uint64_t size, uint8_t* out);

// Decodes `nFrames` frames found in src (host) into a device staging area laid out by frames[].dstOff,
// then copies staging bytes [skip, skip+size) to out (host).
OpStatus host_decode_frames(GpuContext* g, const uint8_t* src, size_t srcSize, const HostFrame* frames, size_t nFrames,
                            uint32_t frameSize, uint64_t skip, uint64_t size, uint8_t* out,
                            const std::function<void()>* whileBusy = nullptr);

// Complete archive from `in` (host) into `out` (host). Mirrors zra::CompressBuffer incl. its metadata quirk.
OpStatus host_compress_buffer(GpuContext* g, const uint8_t* in, size_t n, uint8_t* out, size_t outCap, size_t* written, int level,
                              uint32_t frameSize, bool checksum, const uint8_t* meta, size_t metaSize);

// Frames only (streaming compressor): compresses ceil(n/frameSize) frames back to back into out,
// sizes[i] = compressed size of frame i, *produced = total bytes.
OpStatus host_compress_frames(GpuContext* g, const uint8_t* in, size_t n, uint32_t frameSize, int level, bool checksum,
                              uint8_t* out, size_t outCap, uint64_t* sizes, size_t* produced);

}  // namespace zrab
