// ra_launch.h — host-visible launch interface of ra_kernels.cu (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zrab {

// A batch of random-access requests, all arrays DEVICE pointers. Request i reads sizes[i] bytes
// (uniformSize when sizes == nullptr) at uncompressed offset offsets[i] and writes them to
// out + outOffsets[i] (out + i * uniformSize when outOffsets == nullptr).
struct RaBatch {
  const uint64_t* offsets;
  const uint32_t* sizes;
  const uint64_t* outOffsets;
  uint32_t uniformSize;
  uint32_t frameSize;
  uint64_t uncompressedSize;
};

void launch_ra_map(const RaBatch& b, uint64_t first, uint32_t n, uint32_t* slotOf, uint32_t* uniqueFrames, uint32_t* counters,
                   cudaStream_t st);
void launch_ra_descs(const void* archive, uint64_t tableOff, uint64_t headerSize, uint64_t archiveSize, uint64_t uncompressedSize,
                     uint32_t frameSize, const uint32_t* uniqueFrames, uint32_t nUnique, void* descs, cudaStream_t st);
void launch_ra_gather(const RaBatch& b, uint64_t first, uint32_t n, const uint32_t* slotOf, const void* slots, void* out,
                      cudaStream_t st);
void launch_ra_reset(uint32_t* slotOf, const uint32_t* uniqueFrames, uint32_t nUnique, cudaStream_t st);

}  // namespace zrab
