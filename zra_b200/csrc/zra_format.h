// zra_format.h — the ZRA on-disk layout, serialised field by field (host side).
//
// Layout (all little-endian; reference: source/zra.cpp:96-134, README.md:14-19):
//   @0  u32 0x184D2A50   zstd skippable-frame magic, so stock zstd decoders skip the header
//   @4  u32 headerSize   bytes after this field = 38 + meta + 5*table - 8
//   @8  u32 0x3041525A   "ZRA0"
//   @12 u16 version      1
//   @14 u32 hash         CRC-32 of bytes [0,14) + [18,38) + meta + table
//   @18 u64 uncompressedSize
//   @26 u32 tableSize    frames + 1
//   @30 u32 frameSize
//   @34 u32 metaSize
//   @38 meta[metaSize], then tableSize entries of 5 bytes: 40-bit offset of frame i relative to the
//       first byte after the header; the last entry is the total compressed size.
// The reference gets this layout from #pragma pack + a GCC-only scalar_storage_order pragma; here
// every field is written explicitly so the bytes do not depend on the compiler.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace zrab {

constexpr uint32_t kZraSkippableMagic = 0x184D2A50u;
constexpr uint32_t kZraMagic = 0x3041525Au;
constexpr uint16_t kZraVersion = 1;
constexpr size_t kFixedHeaderSize = 38;
constexpr size_t kEntrySize = 5;
constexpr uint64_t kMaxCompressedSize = 1ull << 40;

inline void put_le(uint8_t* p, uint64_t v, int n) {
  for (int i = 0; i < n; i++) p[i] = static_cast<uint8_t>(v >> (8 * i));
}
inline uint64_t get_le(const uint8_t* p, int n) {
  uint64_t v = 0;
  for (int i = 0; i < n; i++) v |= static_cast<uint64_t>(p[i]) << (8 * i);
  return v;
}

struct FixedHeaderFields {
  uint32_t frameId, headerSize, magic;
  uint16_t version;
  uint32_t hash;
  uint64_t uncompressedSize;
  uint32_t tableSize, frameSize, metaSize;
};

inline FixedHeaderFields parse_fixed_header(const uint8_t* p) {
  FixedHeaderFields f;
  f.frameId = (uint32_t)get_le(p, 4);
  f.headerSize = (uint32_t)get_le(p + 4, 4);
  f.magic = (uint32_t)get_le(p + 8, 4);
  f.version = (uint16_t)get_le(p + 12, 2);
  f.hash = (uint32_t)get_le(p + 14, 4);
  f.uncompressedSize = get_le(p + 18, 8);
  f.tableSize = (uint32_t)get_le(p + 26, 4);
  f.frameSize = (uint32_t)get_le(p + 30, 4);
  f.metaSize = (uint32_t)get_le(p + 34, 4);
  return f;
}

// Writes the 38 fixed bytes with hash = 0.
inline void write_fixed_header(uint8_t* p, uint64_t uncompressedSize, uint32_t tableSize, uint32_t frameSize, uint32_t metaSize) {
  put_le(p, kZraSkippableMagic, 4);
  put_le(p + 4, (uint32_t)(kFixedHeaderSize + metaSize + kEntrySize * (uint64_t)tableSize - 8), 4);
  put_le(p + 8, kZraMagic, 4);
  put_le(p + 12, kZraVersion, 2);
  put_le(p + 14, 0, 4);
  put_le(p + 18, uncompressedSize, 8);
  put_le(p + 26, tableSize, 4);
  put_le(p + 30, frameSize, 4);
  put_le(p + 34, metaSize, 4);
}

// Number of seek-table entries for `size` bytes in `frameSize` frames (source/zra.cpp:190,195,304).
inline uint32_t table_entries(uint64_t size, uint32_t frameSize) {
  return (uint32_t)(size / frameSize + ((size % frameSize) ? 2 : 1));
}

// ZSTD_COMPRESSBOUND (submodule/zstd/lib/zstd.h:174).
inline uint64_t zstd_compress_bound(uint64_t s) {
  return s + (s >> 8) + ((s < (128ull << 10)) ? (((128ull << 10) - s) >> 11) : 0);
}

// CRC-32 (reflected 0xEDB88320, init/xorout all ones) — host copy used for tiny headers; the bulk
// seek-table CRC runs on the device (crc32 kernels). `prev` chains like CRCpp (CRC.h:454-462).
inline uint32_t crc32_host(const uint8_t* p, size_t n, uint32_t prev = 0) {
  uint32_t c = ~prev;
  for (size_t i = 0; i < n; i++) {
    c ^= p[i];
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
  }
  return ~c;
}

// Header hash over a complete header image (fixed + meta + table).
inline uint32_t header_hash_host(const uint8_t* header, size_t total) {
  uint32_t c = crc32_host(header, 14);
  c = crc32_host(header + 18, 20, c);
  return crc32_host(header + kFixedHeaderSize, total - kFixedHeaderSize, c);
}

}  // namespace zrab
