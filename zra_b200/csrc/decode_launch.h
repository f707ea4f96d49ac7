// decode_launch.h — host-visible launch interface of decode_kernels.cu (internal to the library).
#pragma once
#include <cstdio>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace zrab {

// Optional per-kernel timing (bench.py's roofline leg): an event is recorded after every launch
// and the gaps are attributed to the kernel that was just launched. Off by default.
enum KernelId : int { K_START = 0, K_BUILD_DESCS, K_BLOCK_SETUP, K_HUF_DECODE, K_SEQ_DECODE, K_SEQ_EXECUTE, K_FRAME_FINISH,
                      K_COUNT };
const char* kernel_name(int id);

class KernelTimer {
 public:
  ~KernelTimer();
  void mark(int id, cudaStream_t st);   // records an event tagged with the kernel just launched
  void collect();                       // call after a stream synchronize; folds the marks into totals
  void reset();
  void dump_timeline(FILE* f);            // every mark's time relative to the first one (tuning aid, ZRA_B200_TIMELINE=1)
  double ms[K_COUNT] = {};
  uint64_t launches[K_COUNT] = {};

 private:
  struct Mark { int id; cudaEvent_t ev; };
  Mark* marks_{nullptr};
  size_t used_{0}, cap_{0};
  cudaEvent_t* pool_{nullptr};
  size_t poolCap_{0};
};

// Byte offsets of the per-batch device scratch regions, all 256-byte aligned.
struct DecodeLayout {
  size_t offDescs, offCtxs, offTabs, offLit, offSeqs, offSummary, offWork, offHufList, offSeqList, offRedoList;
  uint32_t litStride;  // bytes of literal scratch per frame
  uint32_t seqStride;  // packed sequence records per frame
};

// Scratch needed to decode `nFrames` frames whose largest output is `maxDstCap` bytes.
size_t decode_scratch_bytes(uint32_t nFrames, uint32_t maxDstCap, DecodeLayout* lay);

// summary words (device, at offSummary): [0] lowest failing frame (0xFFFFFFFF = none),
// [1] frames with blocks left to decode, [2] lowest frame with an inconsistent seek-table entry.
void launch_summary_reset(void* scratch, const DecodeLayout& lay, cudaStream_t st);

// Fills the descriptor region from a ZRA seek table that is resident on the device.
void launch_build_descs(const void* archive, uint64_t tableOff, uint64_t headerSize, uint64_t archiveSize,
                        uint64_t uncompressedSize, uint32_t frameSize, uint32_t firstFrame, uint32_t nFrames, uint64_t dstBase,
                        void* scratch, const DecodeLayout& lay, cudaStream_t st, KernelTimer* timer = nullptr);

// A second stream (and two events) on which a round's Huffman stage runs BESIDE its sequence stage: both depend only on
// the block setup, both are latency-bound, and the execute kernel needs both.
struct SideLane {
  cudaStream_t st{nullptr};
  cudaEvent_t fork{nullptr}, join{nullptr};
};

// Runs `rounds` rounds (one block of every frame per round). `first` resets the frame contexts. Returns the number of
// kernels launched.
uint32_t launch_decode_rounds(const void* src, void* dst, uint32_t nFrames, uint32_t rounds, bool first, void* scratch,
                          const DecodeLayout& lay, cudaStream_t st, KernelTimer* timer = nullptr, const SideLane* side = nullptr);

// Checksums + final checks + summary. May be called again after extra rounds.
void launch_frame_finish(const void* src, const void* dst, uint32_t nFrames, void* scratch, const DecodeLayout& lay,
                         cudaStream_t st, KernelTimer* timer = nullptr);

uint32_t frame_status_offset();
uint32_t frame_ctx_size();
uint32_t frame_desc_size();

}  // namespace zrab
