// xxh64.cuh — XXH64 (seed 0) as zstd uses it for the frame content checksum.
// Reference: zstd/lib/common/xxhash.c:415-500 (round/merge/avalanche), :567-720 (digest);
// the frame stores the low 32 bits little-endian (zstd/lib/decompress/zstd_decompress.c:678-686).
//
// Device mapping: the four accumulators of XXH64 are independent until the final merge, so a
// frame is hashed by FOUR lanes (one accumulator each, 8 bytes of every 32-byte stripe); eight
// frames share a warp. Lane 0 of each quad merges, eats the <32-byte tail and avalanches.
#pragma once
#include "zfmt.cuh"

namespace zrab {

constexpr u64 kXP1 = 11400714785074694791ULL;
constexpr u64 kXP2 = 14029467366897019727ULL;
constexpr u64 kXP3 = 1609587929392839161ULL;
constexpr u64 kXP4 = 9650029242287828579ULL;
constexpr u64 kXP5 = 2870177450012600261ULL;

ZRA_DEV u64 xxh_rotl(u64 x, u32 r) { return (x << r) | (x >> (64 - r)); }
ZRA_DEV u64 xxh_round(u64 acc, u64 in) { return xxh_rotl(acc + in * kXP2, 31) * kXP1; }
ZRA_DEV u64 xxh_merge(u64 acc, u64 v) { return (acc ^ xxh_round(0, v)) * kXP1 + kXP4; }
ZRA_DEV u64 xxh_init_acc(u32 lane) {
  return lane == 0 ? kXP1 + kXP2 : (lane == 1 ? kXP2 : (lane == 2 ? 0ull : 0ull - kXP1));
}

// Finishes a hash: `h` is the merged accumulator state (or P5 for short inputs) before adding
// the length; p[0..tail) are the bytes after the last full stripe.
ZRA_DEV u64 xxh_finish(u64 h, u64 totalLen, const u8* p, u32 tail) {
  h += totalLen;
  u32 i = 0;
  for (; i + 8 <= tail; i += 8) { h ^= xxh_round(0, ld64(p + i)); h = xxh_rotl(h, 27) * kXP1 + kXP4; }
  if (i + 4 <= tail) { h ^= (u64)ld32(p + i) * kXP1; h = xxh_rotl(h, 23) * kXP2 + kXP3; i += 4; }
  for (; i < tail; i++) { h ^= p[i] * kXP5; h = xxh_rotl(h, 11) * kXP1; }
  h ^= h >> 33; h *= kXP2; h ^= h >> 29; h *= kXP3; h ^= h >> 32;
  return h;
}

}  // namespace zrab
