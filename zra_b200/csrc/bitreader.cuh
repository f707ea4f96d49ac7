// bitreader.cuh — backward bitstream reader for zstd's FSE / Huffman streams.
//
// zstd writes these streams forward but reads them from the last byte towards the first
// (reference: zstd/lib/common/bitstream.h:272-440; format doc "Bitstreams are read backward").
// The last byte carries a 1-bit end mark above the final padding.
//
// Device design: the whole compressed archive sits in one HBM buffer whose base is 4-byte
// aligned, and a stream is addressed by its absolute byte offset in it. The reader keeps a
// LEFT-JUSTIFIED 64-bit window in two 32-bit registers (hi:lo; the next unread bit is bit 31 of
// hi) so that every extract / consume / refill is one or two funnel-shift instructions (SHF) —
// no 64-bit shifts, which cost several instructions each on the GPU. Words are fetched ALIGNED,
// 32 bits per refill, with the following word always loaded one refill ahead and the following
// 128-byte line prefetched one line ahead: lanes of a warp walk 32 unrelated streams in
// lock-step, so one lane's cache miss would stall all of them.
// Bits below the stream's first byte read as zero — zstd's "stream ran dry" semantics that the
// FSE-compressed Huffman weight decoder relies on; `remaining` goes negative on over-read.
#pragma once
#include "zfmt.cuh"

namespace zrab {

// high 32 bits of (hi:lo) << min(n, 32)
ZRA_DEV u32 fsh_lc(u32 lo, u32 hi, u32 n) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_lc(lo, hi, n);
#else
  return n >= 32 ? lo : (n ? ((hi << n) | (lo >> (32 - n))) : hi);
#endif
}
// low 32 bits of (hi:lo) >> min(n, 32)
ZRA_DEV u32 fsh_rc(u32 lo, u32 hi, u32 n) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_rc(lo, hi, n);
#else
  return n >= 32 ? hi : (n ? ((lo >> n) | (hi << (32 - n))) : lo);
#endif
}

struct BackReader {
  const u32* p;    // next word to load
  u32 hi, lo;      // window; valid bits are the top `avail` bits of hi:lo
  u32 next;        // the word after the window, load already issued (raw)
  u32 nextKeep;    // mask of the bits of `next` that lie inside the stream
  i32 avail;       // 33..64 after refill()
  i32 remaining;   // unread bits of the stream; negative once over-read
  i32 cnt;         // words still to load, the stream's first (partial) word included
  u32 startShift;  // bit offset of the stream's first byte inside its word (0, 8, 16, 24)

  // Issues the load of the next word WITHOUT touching its value (so the warp does not wait for it
  // here); nextKeep says which of its bits belong to the stream and is applied when the word is
  // merged into the window one refill later.
  ZRA_DEV void load_next() {
    u32 v = 0;
    if (cnt > 0) {
#if defined(__CUDA_ARCH__)
      v = __ldg(p);
#else
      v = *p;
#endif
    }
    nextKeep = cnt > 1 ? 0xFFFFFFFFu : (cnt == 1 ? (0xFFFFFFFFu << startShift) : 0u);
    next = v;
    cnt--;
    p--;
  }
  ZRA_DEV u32 take_next() {
    u32 v = next & nextKeep;
    load_next();
    return v;
  }

  ZRA_DEV void prefetch_below() const {
#if defined(__CUDA_ARCH__)
    if (cnt > 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(p - 32));
#endif
  }

  // Returns false if the stream is empty or has no end mark.
  ZRA_DEV bool init(const u8* base, u64 byteOff, u32 len) {
    if (len == 0) return false;
    u32 last = base[byteOff + len - 1];
    if (last == 0) return false;
    u32 total = (len - 1) * 8 + highbit32(last);
    u64 startBit = byteOff * 8;
    u64 endBit = startBit + total;
    i64 ws = (i64)(byteOff >> 2);
    i64 k = endBit ? (i64)((endBit - 1) >> 5) : -1;
    if (k < ws - 1) k = ws - 1;
    startShift = (u32)(byteOff & 3) * 8;
    remaining = (i32)total;
    cnt = (i32)(k - ws + 1);
    p = reinterpret_cast<const u32*>(base) + k;
    u32 top = (u32)(endBit - 32 * (u64)k);  // valid bits in word k (1..32; 32 - 0 when the stream is empty)
    load_next();
    u32 w0 = take_next();
    u32 w1 = take_next();
    hi = fsh_lc(w1, w0, 32 - top);
    lo = fsh_lc(0, w1, 32 - top);
    avail = (i32)top + 32;
    prefetch_below();
    return true;
  }

  // Guarantees at least 33 readable bits in the window.
  ZRA_DEV void refill() {
    if (avail <= 32) {
      u32 v = take_next();
      hi |= fsh_rc(v, 0, (u32)avail);
      lo = fsh_rc(0, v, (u32)avail);
      avail += 32;
#if defined(__CUDA_ARCH__)
      if (((u32)(uintptr_t)p & 127u) == 124u) prefetch_below();
#endif
    }
  }

  // n <= 32, and n <= avail.
  ZRA_DEV u32 peek(u32 n) const { return fsh_lc(hi, 0, n); }
  ZRA_DEV void skip(u32 n) {
    hi = fsh_lc(lo, hi, n);
    lo = fsh_lc(0, lo, n);
    avail -= (i32)n;
    remaining -= (i32)n;
  }
  ZRA_DEV u32 read(u32 n) {
    u32 v = peek(n);
    skip(n);
    return v;
  }
};

}  // namespace zrab
