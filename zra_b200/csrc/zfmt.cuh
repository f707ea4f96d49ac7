// zfmt.cuh — zstd wire-format constants shared by the decode and encode kernels.
//
// Everything here is plain data + tiny pure functions, usable from device code and (for the
// host-side logic tests in tests/host_sim) from ordinary C++: ZRA_DEV expands to
// __host__ __device__ under nvcc and to nothing under g++.
//
// Reference for the numbers: zstd/doc/zstd_compression_format.md (sequence codes :678-765,
// default distributions :835-896) and zstd/lib/common/zstd_internal.h:142-224.
#pragma once
#include <stdint.h>

// ZRA_DEV marks code that runs on the GPU in the product. The same source is compiled as plain
// inline C++ by g++ for tests/host_sim (logic tests without a GPU); it is never a product CPU path.
#if defined(__CUDACC__)
#define ZRA_DEV __device__ __forceinline__
#define ZRA_DEV_NOINLINE __device__ __noinline__
#define ZRA_CONST_TABLE __device__ const
#else
#define ZRA_DEV inline
#define ZRA_DEV_NOINLINE inline
#define ZRA_CONST_TABLE static const
#endif

// Dynamic shared memory of a kernel (the host logic build takes it from the SIMT emulator, tests/host_sim/simt.h).
#if defined(__CUDACC__)
#define ZRA_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#else
#define ZRA_DYN_SMEM(name) unsigned char* name = simt::dyn_smem()
#endif

namespace zrab {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

// ZSTD_ErrorCode values the kernels report (zstd/lib/common/zstd_errors.h:52-79).
enum ZErr : u32 {
  ZE_OK = 0,
  ZE_GENERIC = 1,
  ZE_PREFIX_UNKNOWN = 10,
  ZE_FRAMEPARAM_UNSUPPORTED = 14,
  ZE_WINDOW_TOO_LARGE = 16,
  ZE_CORRUPTION = 20,
  ZE_CHECKSUM_WRONG = 22,
  ZE_DICT_CORRUPTED = 30,
  ZE_DICT_WRONG = 32,
  ZE_TABLELOG_TOO_LARGE = 44,
  ZE_MAXSYM_TOO_SMALL = 48,
  ZE_DST_TOO_SMALL = 70,
  ZE_SRC_WRONG = 72,
};

constexpr u32 kZstdMagic = 0xFD2FB528u;
constexpr u32 kSkippableMagicBase = 0x184D2A50u;
constexpr u32 kBlockSizeMax = 1u << 17;
constexpr u32 kLongNbSeq = 0x7F00;
constexpr u32 kMaxLL = 35, kMaxML = 52, kMaxOF = 31, kDefaultMaxOF = 28;
constexpr u32 kLLFSELog = 9, kMLFSELog = 9, kOFFSELog = 8;
constexpr u32 kLLDefLog = 6, kMLDefLog = 6, kOFDefLog = 5;
constexpr u32 kHufLogMax = 12;

// Literal-length / match-length code → (baseline, extra bits).
ZRA_CONST_TABLE u32 kLLBase[36] = {0,  1,  2,  3,  4,  5,  6,  7,  8,    9,    10,   11,   12,   13,   14,    15,    16,    18,
                                   20, 22, 24, 28, 32, 40, 48, 64, 0x80, 0x100, 0x200, 0x400, 0x800, 0x1000, 0x2000, 0x4000, 0x8000, 0x10000};
ZRA_CONST_TABLE u8 kLLBits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
ZRA_CONST_TABLE u32 kMLBase[53] = {3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18,   19,    20,    21,    22,    23,     24,     25,     26,     27,     28,     29,
                                   30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 0x83, 0x103, 0x203, 0x403, 0x803, 0x1003, 0x2003, 0x4003, 0x8003, 0x10003};
ZRA_CONST_TABLE u8 kMLBits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                  0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
// Default (predefined) normalised distributions.
ZRA_CONST_TABLE int16_t kLLDefNorm[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
ZRA_CONST_TABLE int16_t kMLDefNorm[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                          1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
ZRA_CONST_TABLE int16_t kOFDefNorm[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

// Offset code c: extra bits = c, baseline = 2^c - 3 for c >= 2 (values 0/1 for the repcode codes).
ZRA_DEV u32 of_base(u32 c) { return c < 2 ? c : ((1u << c) - 3u); }

ZRA_DEV u32 highbit32(u32 v) {
#if defined(__CUDA_ARCH__)
  return 31u - (u32)__clz((int)v);
#else
  return 31u - (u32)__builtin_clz(v);
#endif
}

// Little-endian reads from byte pointers of unknown alignment.
ZRA_DEV u32 ld16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
ZRA_DEV u32 ld24(const u8* p) { return ld16(p) | ((u32)p[2] << 16); }
ZRA_DEV u32 ld32(const u8* p) { return ld16(p) | (ld16(p + 2) << 16); }
ZRA_DEV u64 ld64(const u8* p) { return (u64)ld32(p) | ((u64)ld32(p + 4) << 32); }

}  // namespace zrab
