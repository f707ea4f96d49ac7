// gpu_context.h — the library's GPU context (internal). One per host thread / per C-ABI handle.
//
// Owns: the CUDA device binding, a private stream, grow-only device buffers (decode scratch,
// staged input, staged output) and pinned host staging. zra::ZDCtx / zra::ZCCtx — opaque in the
// reference, where they wrap ZSTD_DCtx / ZSTD_CCtx (source/zra.cpp:25-44) — wrap this class.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "decode_launch.h"
#include "ra_launch.h"

namespace zrab {

struct DevBuf {
  void* p{nullptr};
  size_t cap{0};
};

struct HostFrame {  // same layout as ZraCudaFrame / zrab::FrameDesc
  uint64_t srcOff, dstOff;
  uint32_t srcLen, dstCap, exact, pad;
};

struct DecodeResult {
  int zstd{0};                  // ZSTD_ErrorCode of the lowest failing frame, 0 = ok
  uint32_t failedFrame{~0u};    // its index
  bool cudaFailed{false};
};

// Optional host<->device staging folded into the chunk pipeline (needs a host `frames` array).
struct HostStaging {
  const uint8_t* hostSrc{nullptr};  // host copy of the source buffer, same offsets as dSrc: chunk ranges are uploaded
  uint8_t* hostDst{nullptr};        // receives device output bytes [dstSkip, dstSkip + dstSize)
  uint64_t dstSkip{0};
  uint64_t dstSize{~0ull};
  // called once on the caller's thread after the first group's work is queued and before it is waited for (the streaming
  // classes read the NEXT call's bytes through the user's callback here, beside the GPU work of this call)
  const std::function<void()>* whileBusy{nullptr};
};

// Parsed fixed header of a ZRA archive (source/zra.cpp:111-134 layout).
struct ArchiveInfo {
  uint32_t headerSize;  // fixed + meta + table
  uint32_t tableSize, frameSize, metaSize;
  uint64_t uncompressedSize;
  uint64_t frames;
};

class GpuContext {
 public:
  explicit GpuContext(int device);
  ~GpuContext();
  GpuContext(const GpuContext&) = delete;
  GpuContext& operator=(const GpuContext&) = delete;

  bool ok() const { return ok_; }
  const std::string& last_error() const { return lastError_; }
  uint64_t launches() const { return launches_; }
  cudaStream_t stream() const { return stream_; }
  int device() const { return device_; }

  // Records a CUDA failure (returns true if `e` is an error).
  bool check(cudaError_t e, const char* what);
  void fail(const std::string& msg) { lastError_ = msg; }

  // Grow-only device buffers.
  void* ensure(DevBuf& b, size_t bytes);
  // Page-locked host staging that the streaming classes hand to the user's read callback, so that the compressed
  // bytes land where the upload can start from at link speed (grown geometrically, kept for the context's lifetime).
  uint8_t* pinned_stage(size_t bytes, int which = 0);  // two buffers: this call's bytes and the read-ahead of the next
  // bumped every time a buffer is handed out: a reader that parked bytes in one can tell whether anybody asked for it since
  uint64_t stage_gen(int which) const { return stageGen_[which & 1]; }
  DevBuf scratch, stageIn, stageOut, misc;
  DevBuf raSlotOf, raUnique, raDescs, raFrames;  // batched random access (ra_context.cu)

  // Decodes frames described either by a host array (`frames`) or, when frames == nullptr, by the
  // seek table of a device-resident archive (`info`, firstFrame, dstBase). Synchronous.
  // The frames are cut into chunks that run on a pool of streams: the phases of different chunks
  // (entropy decode is latency-bound, sequence execution is issue-bound) overlap on the SMs, and
  // with `io` the PCIe copies of one chunk overlap the kernels of the others.
  DecodeResult decode(const void* dSrc, size_t srcSize, const HostFrame* frames, const ArchiveInfo* info, uint64_t firstFrame,
                      uint64_t nFrames, uint32_t maxDstCap, void* dDst, uint32_t* frameSizes, cudaStream_t st,
                      const HostStaging* io = nullptr);

  // Batched random access over a device-resident archive: every frame a
  // batch touches is decoded once, the requested slices are gathered into dOut. All arrays of `b`
  // are device pointers; maxSize bounds every request's size. Synchronous.
  struct RaResult {
    bool cudaFailed{false};
    int zra{0}, zstd{0};          // zra::StatusCode / ZSTD_ErrorCode
    uint64_t badRequest{~0ull};   // first request that is out of bounds (zra == 5)
    uint64_t uniqueFrames{0};     // frames decoded (after de-duplication, summed over sub-batches)
  };
  RaResult random_access(const void* dArchive, size_t archiveSize, const ArchiveInfo& info, RaBatch b, uint64_t count,
                         uint32_t maxSize, void* dOut, cudaStream_t st);

  // Compression results.
  struct CompressStatus {
    bool cudaFailed{false};
    int zra{0};           // zra::StatusCode value (0 ok, 6 output too small, 7 compressed size too large)
    uint64_t total{0};    // bytes produced
  };
  // dIn[0..n) -> complete archive at dOut (header, metadata, 40-bit seek table, frames), all on the device.
  // refMetaQuirk reproduces zra::CompressBuffer's layout for a non-empty meta: the
  // header records metaSize but no metadata bytes are stored.
  // hostIn / hostOut (both or neither): the input lives in HOST memory and the archive is wanted there. The frames are
  // then compressed in batches whose upload (batch b + 1), kernels (batch b) and download (batch b - 1) overlap.
  // hostFrames (optional): the frames go there instead of behind the header at hostOut (hostOut may then be null: the
  // header and the seek table stay on the device, at dOut).
  CompressStatus compress_archive(const void* dIn, size_t n, void* dOut, size_t outCap, int level, uint32_t frameSize,
                                  bool checksum, const uint8_t* metaHost, size_t metaSize, bool refMetaQuirk, cudaStream_t st,
                                  const uint8_t* hostIn = nullptr, uint8_t* hostOut = nullptr, uint8_t* hostFrames = nullptr);
  // dIn[0..n) -> zstd frames back to back at dOut (no header); sizesHost[i] = compressed size of frame i.
  CompressStatus compress_frames(const void* dIn, size_t n, uint32_t frameSize, int level, bool checksum, void* dOut,
                                 size_t outCap, uint64_t* sizesHost, cudaStream_t st);

  void bind();  // cudaSetDevice(device_)

  // Per-kernel event timing (off by default; adds an event record per launch when on).
  void set_profiling(bool on) { profiling_ = on; }
  KernelTimer timer;

 private:
  int device_{0};
  bool ok_{false};
  cudaStream_t stream_{nullptr};
  std::string lastError_;
  uint64_t launches_{0};
  bool profiling_{false};
  static constexpr int kPoolStreams = 16;
  cudaStream_t pool_[kPoolStreams] = {};
  SideLane side_[kPoolStreams];  // the chunks' Huffman stages run beside their sequence stages (decode_launch.h)
  cudaEvent_t forkEvent_{nullptr};
  // host-pointer calls: every upload goes through upStream_ and every download through downStream_, in chunk order
  // (copies issued on the chunks' own streams share the link and all finish late, which delays the first download)
  cudaStream_t upStream_{nullptr}, downStream_{nullptr};
  std::vector<cudaEvent_t> upEvents_, doneEvents_;
  bool ensure_events(size_t n);
  uint32_t* summaryHost_{nullptr};  // pinned, 4 words per chunk
  uint8_t* pinnedStage_[2] = {nullptr, nullptr};
  size_t pinnedStageCap_[2] = {0, 0};
  uint64_t stageGen_[2] = {0, 0};
  uint32_t* raHost_{nullptr};       // pinned, {unique frames, first bad request}
  uint64_t raSlotFrames_{0};        // entries of raSlotOf that are initialised to "empty"
  static constexpr uint32_t kMaxChunks = 1024;
};

// Thread-local default context used by the zra.h / zra.hpp host-pointer entry points.
GpuContext* default_context();

}  // namespace zrab
