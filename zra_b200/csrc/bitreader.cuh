// bitreader.cuh — backward bitstream reader for zstd's FSE / Huffman streams.
//
// zstd writes these streams forward but reads them from the last byte towards the first
// (reference: zstd/lib/common/bitstream.h:272-440; format doc "Bitstreams are read backward").
// The last byte carries a 1-bit end mark above the final padding.
//
// Device design: the whole compressed archive sits in one HBM buffer whose base is 4-byte
// aligned. A stream is addressed by its absolute byte offset in that buffer; the reader keeps a
// 64-bit window in registers and refills it 32 bits at a time from ALIGNED words, with the next
// word always prefetched one refill ahead so the (L1/L2) load latency overlaps decoding.
// Bits below the stream's first byte are forced to zero — zstd's "stream ran dry" semantics that
// the FSE-compressed Huffman weight decoder relies on.
#pragma once
#include "zfmt.cuh"

namespace zrab {

struct BackReader {
  const u32* words;  // aligned base of the source buffer
  u64 w;             // window: the unread bits are w[0 .. pos)
  u32 next;          // word nextIdx, already loaded
  i32 pos;           // unread bits held in w (33..64 after refill())
  i64 nextIdx;       // index of `next`
  i64 remaining;     // unread bits of the stream; negative once over-read
  u64 startBit;      // absolute bit index of the stream's first bit

  ZRA_DEV u32 load_word(i64 k) const {
    i64 low = k * 32;
    if (low + 32 <= (i64)startBit) return 0;  // entirely below the stream (also covers k < 0)
#if defined(__CUDA_ARCH__)
    u32 v = __ldg(words + k);
#else
    u32 v = words[k];
#endif
    if (low < (i64)startBit) {
      u32 d = (u32)((i64)startBit - low);  // 1..31
      v = (v >> d) << d;
    }
    return v;
  }

  // Returns false if the stream is empty or has no end mark.
  ZRA_DEV bool init(const u8* base, u64 byteOff, u32 len) {
    words = reinterpret_cast<const u32*>(base);
    if (len == 0) return false;
    u32 last = base[byteOff + len - 1];
    if (last == 0) return false;
    u32 total = (len - 1) * 8 + highbit32(last);
    startBit = byteOff * 8;
    u64 endBit = startBit + total;
    remaining = total;
    i64 k = endBit ? (i64)((endBit - 1) >> 5) : 0;
    w = ((u64)load_word(k) << 32) | load_word(k - 1);
    pos = (i32)((i64)endBit - 32 * (k - 1));
    nextIdx = k - 2;
    next = load_word(nextIdx);
    return true;
  }

  // Guarantees at least 33 readable bits in the window.
  ZRA_DEV void refill() {
    if (pos <= 32) {
      w = (w << 32) | next;
      pos += 32;
      nextIdx--;
      next = load_word(nextIdx);
    }
  }

  // n <= 32 and n <= pos.
  // (the & 63 only matters for pos == 64, n == 0, where the mask is 0 anyway)
  ZRA_DEV u32 peek(u32 n) const { return (u32)((w >> ((u32)(pos - (i32)n) & 63u)) & ((1ull << n) - 1ull)); }
  ZRA_DEV void skip(u32 n) {
    pos -= (i32)n;
    remaining -= n;
  }
  ZRA_DEV u32 read(u32 n) {
    u32 v = peek(n);
    skip(n);
    return v;
  }
};

}  // namespace zrab
