// ref_harness.cpp — thin multi-threaded driver AROUND the unmodified reference (test infrastructure).
//
// Linked into oracle/_ref/libzra_ref.so next to the reference's own objects. Nothing here
// re-implements ZRA: every frame is produced / consumed by the reference's zra::Compressor,
// zra::Decompressor and zra::DecompressRA (source/zra.cpp:304-424, 258-296). The harness only
// adds what the reference does not ship: running those objects on T host threads
// over disjoint frame-aligned shards so that (a) big bench archives can be produced in reasonable
// time and (b) an "all host cores" CPU baseline exists beside the faithful single-thread one.
// The stitched header is written by a final zra::Compressor-compatible serialisation whose bytes
// tests/test_oracle.py compares with a serial zra::CompressBuffer archive.
#include <zstd.h>
#include <zstd_errors.h>

#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "zra.hpp"

#define CRCPP_USE_CPP11 1
#include <CRC.h>

#define HARNESS_API extern "C" __attribute__((visibility("default")))

namespace {
  struct Shard {
    size_t firstFrame{}, frames{};
    std::vector<uint8_t> bytes;   // compressed frames of this shard, back to back
    std::vector<uint64_t> sizes;  // compressed size of every frame
    int zra{}, zstd{};
  };

  void put(uint8_t* p, uint64_t v, int n) {
    for (int i = 0; i < n; i++) p[i] = static_cast<uint8_t>(v >> (8 * i));
  }
}  // namespace

// Compresses `in` as a ZRA archive with T threads. Every thread owns one zra::Compressor over a
// contiguous run of frames; tables are stitched afterwards. Returns 0 or a ZRA status code.
HARNESS_API int ref_compress_mt(const void* in, size_t n, void* out, size_t outCap, size_t* outSize, int level,
                                uint32_t frameSize, int checksum, int threads) {
  size_t frames = n / frameSize + ((n % frameSize) ? 1 : 0);
  size_t T = std::max<size_t>(1, std::min<size_t>(threads, frames ? frames : 1));
  std::vector<Shard> shards(T);
  std::vector<std::thread> pool;
  for (size_t t = 0; t < T; t++) {
    shards[t].firstFrame = frames * t / T;
    shards[t].frames = frames * (t + 1) / T - shards[t].firstFrame;
    pool.emplace_back([&, t] {
      Shard& s = shards[t];
      if (!s.frames) return;
      size_t begin = s.firstFrame * static_cast<size_t>(frameSize);
      size_t end = std::min(n, (s.firstFrame + s.frames) * static_cast<size_t>(frameSize));
      try {
        // one frame per call so that the per-frame compressed sizes are observable
        zra::Compressor c(end - begin, static_cast<int8_t>(level), frameSize, checksum != 0);
        zra::Buffer tmp;
        for (size_t off = begin; off < end; off += frameSize) {
          size_t len = std::min<size_t>(frameSize, end - off);
          c.Compress(zra::BufferView(const_cast<uint8_t*>(static_cast<const uint8_t*>(in)) + off, len), tmp);
          s.sizes.push_back(tmp.size());
          s.bytes.insert(s.bytes.end(), tmp.begin(), tmp.end());
        }
      } catch (const zra::Exception& e) {
        s.zra = static_cast<int>(e.code);
        s.zstd = e.zstdCode;
      }
    });
  }
  for (auto& th : pool) th.join();
  for (auto& s : shards)
    if (s.zra) return s.zra;

  uint32_t tableSize = static_cast<uint32_t>(frames + 1);
  size_t headerBytes = 38 + 5ull * tableSize;
  size_t total = headerBytes;
  for (auto& s : shards) total += s.bytes.size();
  *outSize = total;
  if (total > outCap) return static_cast<int>(zra::StatusCode::OutputBufferTooSmall);

  auto* p = static_cast<uint8_t*>(out);
  put(p, 0x184D2A50u, 4);
  put(p + 4, headerBytes - 8, 4);
  put(p + 8, 0x3041525Au, 4);
  put(p + 12, 1, 2);
  put(p + 14, 0, 4);
  put(p + 18, n, 8);
  put(p + 26, tableSize, 4);
  put(p + 30, frameSize, 4);
  put(p + 34, 0, 4);
  uint8_t* entry = p + 38;
  uint8_t* body = p + headerBytes;
  uint64_t running = 0;
  for (auto& s : shards) {
    for (uint64_t sz : s.sizes) {
      put(entry, running, 5);
      entry += 5;
      running += sz;
    }
    std::memcpy(body, s.bytes.data(), s.bytes.size());
    body += s.bytes.size();
  }
  put(entry, running, 5);
  auto crc = CRC::Calculate(p, 14, CRC::CRC_32());
  crc = CRC::Calculate(p + 18, 20, CRC::CRC_32(), crc);
  crc = CRC::Calculate(p + 38, headerBytes - 38, CRC::CRC_32(), crc);
  put(p + 14, crc, 4);
  return 0;
}

// Whole-archive decode with T threads, each driving its own zra::Decompressor (memcpy read callback)
// over a disjoint frame-aligned range. status[0]=zra code, status[1]=zstd code of the first failure.
HARNESS_API int ref_decompress_mt(const void* in, size_t n, void* out, size_t outCap, int threads, int* status) {
  status[0] = status[1] = 0;
  const auto* base = static_cast<const uint8_t*>(in);
  auto reader = [base, n](size_t off, size_t size, void* buf) {
    if (off + size > n) throw zra::Exception(zra::StatusCode::OutOfBoundsAccess);
    std::memcpy(buf, base + off, size);
  };
  try {
    zra::Header h(reader);
    if (outCap < h.uncompressedSize) {
      status[0] = static_cast<int>(zra::StatusCode::OutputBufferTooSmall);
      return -1;
    }
    size_t frames = h.uncompressedSize / h.frameSize + ((h.uncompressedSize % h.frameSize) ? 1 : 0);
    size_t T = std::max<size_t>(1, std::min<size_t>(threads, frames ? frames : 1));
    std::atomic<int> fz{0}, fs{0};
    std::vector<std::thread> pool;
    for (size_t t = 0; t < T; t++) {
      pool.emplace_back([&, t] {
        size_t f0 = frames * t / T, f1 = frames * (t + 1) / T;
        if (f0 == f1) return;
        size_t begin = f0 * static_cast<size_t>(h.frameSize);
        size_t end = std::min<size_t>(h.uncompressedSize, f1 * static_cast<size_t>(h.frameSize));
        try {
          zra::Decompressor d(reader, ~size_t{0} >> 1);
          // 256 frames per call keeps the per-call cache small like a streaming user would
          size_t step = 256ull * h.frameSize;
          for (size_t off = begin; off < end; off += step) {
            size_t len = std::min(step, end - off);
            d.Decompress(off, len, zra::BufferView(static_cast<uint8_t*>(out) + off, len));
          }
        } catch (const zra::Exception& e) {
          fz = static_cast<int>(e.code);
          fs = e.zstdCode;
        }
      });
    }
    for (auto& th : pool) th.join();
    status[0] = fz;
    status[1] = fs;
    return fz ? -1 : 0;
  } catch (const zra::Exception& e) {
    status[0] = static_cast<int>(e.code);
    status[1] = e.zstdCode;
    return -1;
  }
}

// `count` random reads of `size` bytes (offsets[i]) into out[i*size..], T threads, one zra::Decompressor each.
HARNESS_API int ref_ra_mt(const void* in, size_t n, const uint64_t* offsets, size_t count, size_t size, void* out,
                          int threads, int* status) {
  status[0] = status[1] = 0;
  const auto* base = static_cast<const uint8_t*>(in);
  auto reader = [base, n](size_t off, size_t sz, void* buf) {
    if (off + sz > n) throw zra::Exception(zra::StatusCode::OutOfBoundsAccess);
    std::memcpy(buf, base + off, sz);
  };
  size_t T = std::max<size_t>(1, std::min<size_t>(threads, count ? count : 1));
  std::atomic<int> fz{0}, fs{0};
  std::vector<std::thread> pool;
  for (size_t t = 0; t < T; t++) {
    pool.emplace_back([&, t] {
      try {
        zra::Decompressor d(reader);
        for (size_t i = count * t / T; i < count * (t + 1) / T; i++)
          d.Decompress(offsets[i], size, zra::BufferView(static_cast<uint8_t*>(out) + i * size, size));
      } catch (const zra::Exception& e) {
        fz = static_cast<int>(e.code);
        fs = e.zstdCode;
      }
    });
  }
  for (auto& th : pool) th.join();
  status[0] = fz;
  status[1] = fs;
  return fz ? -1 : 0;
}

// The in-memory zra::DecompressRA called in a loop on one thread (the faithful reference RA path).
HARNESS_API int ref_ra_inmemory(const void* in, size_t n, const uint64_t* offsets, size_t count, size_t size, void* out,
                                int* status) {
  status[0] = status[1] = 0;
  try {
    zra::BufferView view(const_cast<void*>(in), n);
    for (size_t i = 0; i < count; i++)
      zra::DecompressRA(view, zra::BufferView(static_cast<uint8_t*>(out) + i * size, size), offsets[i], size);
    return 0;
  } catch (const zra::Exception& e) {
    status[0] = static_cast<int>(e.code);
    status[1] = e.zstdCode;
    return -1;
  }
}

// Plain ZSTD_decompress / ZSTD_compress of the vendored zstd, for frame-level checks.
HARNESS_API long long ref_zstd_decompress(void* dst, size_t cap, const void* src, size_t n) {
  size_t r = ZSTD_decompress(dst, cap, src, n);
  return ZSTD_isError(r) ? -static_cast<long long>(ZSTD_getErrorCode(r)) : static_cast<long long>(r);
}

HARNESS_API unsigned ref_crc32(const void* p, size_t n, unsigned prev, int chained) {
  return chained ? CRC::Calculate(p, n, CRC::CRC_32(), prev) : CRC::Calculate(p, n, CRC::CRC_32());
}

HARNESS_API int ref_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }
