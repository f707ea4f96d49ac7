// entropy.cuh — FSE / Huffman table construction for the decode kernels.
//
// One GPU thread builds the tables of one block (thread-serial code; the parallelism is across
// frames). What each routine follows in the reference (zstd/ = submodule/zstd/lib):
//   fse_read_ncount ........ zstd/common/entropy_common.c:41-145   (FSE_readNCount)
//   fse_build_seq_table .... zstd/decompress/zstd_decompress_block.c:367-427 (ZSTD_buildFSETable)
//   huf_read_weights ....... zstd/common/entropy_common.c:155-216  (HUF_readStats) and
//                            zstd/common/fse_decompress.c:177-273  (FSE_decompress_wksp, 2 states)
//   huf_build_table ........ zstd/decompress/huf_decompress.c:118-185 (HUF_readDTableX1_wksp)
#pragma once
#include "bitreader.cuh"

namespace zrab {

// Sequence decode-table entry, 8 bytes: nextState | nbAddBits<<16 | nbBits<<24 | base<<32.
typedef u64 SeqSym;
ZRA_DEV SeqSym seqsym_pack(u32 nextState, u32 nbAddBits, u32 nbBits, u32 base) {
  return (u64)nextState | ((u64)nbAddBits << 16) | ((u64)nbBits << 24) | ((u64)base << 32);
}
ZRA_DEV u32 seqsym_next(SeqSym e) { return (u32)e & 0xFFFFu; }
ZRA_DEV u32 seqsym_addbits(SeqSym e) { return ((u32)e >> 16) & 0xFFu; }
ZRA_DEV u32 seqsym_nbbits(SeqSym e) { return (u32)e >> 24; }
ZRA_DEV u32 seqsym_base(SeqSym e) { return (u32)(e >> 32); }

enum SeqKind : u32 { SEQ_LL = 0, SEQ_OF = 1, SEQ_ML = 2, SEQ_PLAIN = 3 };

ZRA_DEV u32 seq_base(u32 kind, u32 s) {
  if (kind == SEQ_LL) return kLLBase[s];
  if (kind == SEQ_ML) return kMLBase[s];
  if (kind == SEQ_OF) return of_base(s);
  return s;  // SEQ_PLAIN keeps the symbol itself
}
ZRA_DEV u32 seq_addbits(u32 kind, u32 s) {
  if (kind == SEQ_LL) return kLLBits[s];
  if (kind == SEQ_ML) return kMLBits[s];
  if (kind == SEQ_OF) return s;
  return 0;
}

// Forward LSB-first peek of n (<=16) bits at bit position `pos` of src[0..len); zeros past the end.
ZRA_DEV u32 fwd_peek(const u8* src, u32 len, u32 pos, u32 n) {
  u32 byte = pos >> 3;
  u32 v = 0;
  for (u32 i = 0; i < 4; i++) {
    u32 b = (byte + i < len) ? src[byte + i] : 0;
    v |= b << (8 * i);
  }
  return (v >> (pos & 7)) & ((1u << n) - 1u);
}

// Parses a normalised-count header. Returns bytes consumed (>0) or 0 on error (err set).
ZRA_DEV u32 fse_read_ncount(const u8* src, u32 len, int16_t* norm, u32* maxSymbol, u32* tableLog, u32* err) {
  if (len < 1) { *err = ZE_SRC_WRONG; return 0; }
  u32 pos = 0;
  u32 log = fwd_peek(src, len, pos, 4) + 5;
  pos += 4;
  if (log > 15) { *err = ZE_TABLELOG_TOO_LARGE; return 0; }
  *tableLog = log;
  i32 remaining = (1 << log) + 1;
  i32 threshold = 1 << log;
  u32 nbBits = log + 1;
  u32 sym = 0;
  bool previous0 = false;
  while (remaining > 1 && sym <= *maxSymbol) {
    if (previous0) {
      u32 n0 = sym;
      for (;;) {
        u32 v = fwd_peek(src, len, pos, 2);
        pos += 2;
        n0 += v;
        if (v != 3) break;
        if (pos > len * 8) break;
      }
      if (n0 > *maxSymbol) { *err = ZE_MAXSYM_TOO_SMALL; return 0; }
      while (sym < n0) norm[sym++] = 0;
    }
    i32 max = (2 * threshold - 1) - remaining;
    i32 count;
    u32 v = fwd_peek(src, len, pos, nbBits);
    if ((i32)(v & (u32)(threshold - 1)) < max) {
      count = (i32)(v & (u32)(threshold - 1));
      pos += nbBits - 1;
    } else {
      count = (i32)(v & (u32)(2 * threshold - 1));
      if (count >= threshold) count -= max;
      pos += nbBits;
    }
    count--;
    remaining -= count < 0 ? -count : count;
    norm[sym++] = (int16_t)count;
    previous0 = (count == 0);
    while (remaining < threshold) { nbBits--; threshold >>= 1; }
  }
  if (remaining != 1 || pos > len * 8) { *err = ZE_CORRUPTION; return 0; }
  *maxSymbol = sym - 1;
  return (pos + 7) >> 3;
}

// Builds a decode table of 1<<log entries from normalised counts. `table` may live in global or
// shared memory; the symbol of every cell is parked in the entry's high half during the spread.
ZRA_DEV void fse_build_seq_table(SeqSym* table, const int16_t* norm, u32 maxSymbol, u32 log, u32 kind) {
  u32 size = 1u << log, mask = size - 1, high = size - 1;
  u16 nextv[256];
  for (u32 s = 0; s <= maxSymbol; s++) {
    if (norm[s] == -1) { table[high--] = (u64)s << 32; nextv[s] = 1; }
    else nextv[s] = (u16)norm[s];
  }
  u32 step = (size >> 1) + (size >> 3) + 3, pos = 0;
  for (u32 s = 0; s <= maxSymbol; s++) {
    for (i32 i = 0; i < norm[s]; i++) {
      table[pos] = (u64)s << 32;
      do { pos = (pos + step) & mask; } while (pos > high);
    }
  }
  for (u32 u = 0; u < size; u++) {
    u32 s = (u32)(table[u] >> 32);
    u32 ns = nextv[s]++;
    u32 nb = log - highbit32(ns);
    table[u] = seqsym_pack((ns << nb) - size, seq_addbits(kind, s), nb, seq_base(kind, s));
  }
}

ZRA_DEV void fse_build_rle_table(SeqSym* table, u32 s, u32 kind) { table[0] = seqsym_pack(0, seq_addbits(kind, s), 0, seq_base(kind, s)); }

// ---- compact sequence tables (the format the decode kernels keep in shared memory) ----------
// 2-byte entry: symbol (6 bits) | ns << 6, where ns is the "next state" counter of the cell
// (range [count, 2*count) <= 1023). nbBits and the next-state base are recomputed on the fly:
//   nbBits = log - highbit(ns),  nextStateBase = (ns << nbBits) - (1 << log)
// Halving the entry (the reference's ZSTD_seqSymbol is 8 bytes, zstd_decompress_internal.h:62-82)
// doubles the number of frames whose three tables fit in one SM's shared memory.
typedef u16 CSym;
ZRA_DEV u32 csym_symbol(CSym e) { return e & 63u; }
ZRA_DEV u32 csym_ns(CSym e) { return (u32)e >> 6; }

ZRA_DEV void fse_build_compact(CSym* table, const int16_t* norm, u32 maxSymbol, u32 log) {
  u32 size = 1u << log, mask = size - 1, high = size - 1;
  u16 nextv[64];
  for (u32 s = 0; s <= maxSymbol; s++) {
    if (norm[s] == -1) { table[high--] = (CSym)s; nextv[s] = 1; }
    else nextv[s] = (u16)norm[s];
  }
  u32 step = (size >> 1) + (size >> 3) + 3, pos = 0;
  for (u32 s = 0; s <= maxSymbol; s++) {
    for (i32 i = 0; i < norm[s]; i++) {
      table[pos] = (CSym)s;
      do { pos = (pos + step) & mask; } while (pos > high);
    }
  }
  for (u32 u = 0; u < size; u++) {
    u32 s = table[u];
    u32 ns = nextv[s]++;
    table[u] = (CSym)(s | (ns << 6));
  }
}
ZRA_DEV void fse_build_compact_rle(CSym* table, u32 s) { table[0] = (CSym)(s | (1u << 6)); }

// Packed per-code lookup: baseline (24 bits) | extra bits << 24.
ZRA_DEV u32 ll_lut(u32 code) { return kLLBase[code] | ((u32)kLLBits[code] << 24); }
ZRA_DEV u32 ml_lut(u32 code) { return kMLBase[code] | ((u32)kMLBits[code] << 24); }

// Huffman decode-table entry: symbol | nbBits << 8.
typedef u16 HufSym;

// Reads the weights of a Huffman tree description at src[0..len). On success returns the bytes
// consumed, fills weights[0..*count) (last weight implied) and *tableLog. Returns 0 on error.
ZRA_DEV u32 huf_read_weights(const u8* base, u64 off, u32 len, u8* weights, u32* count, u32* tableLog) {
  const u8* src = base + off;
  if (len < 1) return 0;
  u32 h = src[0];
  u32 n, consumed;
  if (h >= 128) {
    n = h - 127;
    consumed = (n + 1) / 2;
    if (consumed + 1 > len) return 0;
    for (u32 i = 0; i < n; i += 2) {
      weights[i] = src[1 + i / 2] >> 4;
      weights[i + 1] = src[1 + i / 2] & 15;
    }
  } else {
    consumed = h;
    if (consumed + 1 > len || h == 0) return 0;
    int16_t norm[256];
    u32 maxSym = 255, log, err = 0;
    u32 hs = fse_read_ncount(src + 1, h, norm, &maxSym, &log, &err);
    if (!hs || log > 6) return 0;
    SeqSym t[64];
    fse_build_seq_table(t, norm, maxSym, log, SEQ_PLAIN);
    BackReader br;
    if (!br.init(base, off + 1 + hs, h - hs)) return 0;
    u32 s1 = br.read(log);
    u32 s2 = br.read(log);
    br.refill();
    n = 0;
    for (;;) {
      if (n > 253) return 0;
      weights[n++] = (u8)seqsym_base(t[s1]);
      s1 = seqsym_next(t[s1]) + br.read(seqsym_nbbits(t[s1]));
      br.refill();
      if (br.remaining < 0) { weights[n++] = (u8)seqsym_base(t[s2]); break; }
      if (n > 253) return 0;
      weights[n++] = (u8)seqsym_base(t[s2]);
      s2 = seqsym_next(t[s2]) + br.read(seqsym_nbbits(t[s2]));
      br.refill();
      if (br.remaining < 0) { weights[n++] = (u8)seqsym_base(t[s1]); break; }
    }
  }
  u32 total = 0;
  for (u32 i = 0; i < n; i++) {
    if (weights[i] >= kHufLogMax) return 0;
    total += (1u << weights[i]) >> 1;
  }
  if (total == 0) return 0;
  u32 log = highbit32(total) + 1;
  if (log > kHufLogMax) return 0;
  u32 rest = (1u << log) - total;
  u32 hb = highbit32(rest);
  if ((1u << hb) != rest) return 0;
  weights[n++] = (u8)(hb + 1);
  *count = n;
  *tableLog = log;
  return consumed + 1;
}

// Two-level form of the same single-symbol table, small enough that every frame in flight keeps
// its table in shared memory (a 2^11-entry table is 4 KiB; this is typically < 0.8 KiB):
//   L1  2^p entries, p = min(log, 8), indexed by the next p bits. Codes of at most p bits resolve here.
//   L2  the first l2 entries of the FULL 2^log table, indexed by the next log bits. zstd fills the
//       table by ascending weight, i.e. the codes LONGER than p bits occupy exactly its first l2
//       entries, and l2 is a multiple of 2^(log-p): L1 cells below nEsc = l2 >> (log-p) escape to L2.
// Layout in `tab`: L1 at [0, 2^p), L2 at [2^p, 2^p + l2). Returns false when cap entries do not hold it.
struct HufLevels {
  u32 log, p, nEsc, l2;
};
ZRA_DEV bool huf_build_two_level(HufSym* tab, u32 cap, const u8* weights, u32 count, u32 log, HufLevels* lv) {
  const u32 p = log < 8 ? log : 8;
  const u32 drop = log - p;  // a code of weight w has log+1-w bits: longer than p <=> w <= drop
  u32 rank[16];
  for (u32 r = 0; r < 16; r++) rank[r] = 0;
  for (u32 i = 0; i < count; i++) rank[weights[i]]++;
  u32 start[16], nxt = 0;
  for (u32 r = 1; r <= log; r++) {
    start[r] = nxt;
    nxt += rank[r] << (r - 1);
  }
  const u32 l2 = drop ? start[drop + 1] : 0;
  lv->log = log; lv->p = p; lv->l2 = l2; lv->nEsc = l2 >> drop;
  if ((1u << p) + l2 > cap) return false;
  HufSym* l1 = tab;
  HufSym* t2 = tab + (1u << p);
  for (u32 s = 0; s < count; s++) {
    u32 wt = weights[s];
    if (!wt) continue;
    u32 span = (1u << wt) >> 1;
    HufSym e = (HufSym)(s | ((log + 1 - wt) << 8));
    u32 b = start[wt];
    start[wt] = b + span;
    if (wt <= drop) {
      for (u32 u = 0; u < span; u++) t2[b + u] = e;
    } else {
      u32 c0 = b >> drop, cn = span >> drop;
      for (u32 u = 0; u < cn; u++) l1[c0 + u] = e;
    }
  }
  return true;
}
// One symbol off the top of a left-justified 32-bit window `hi`: returns the table entry.
ZRA_DEV u32 huf_lookup(const HufSym* tab, const HufLevels& lv, u32 hi) {
  const u32 i8 = fsh_lc(hi, 0, lv.p);
  const u32 iL = fsh_lc(hi, 0, lv.log);
  const u32 idx = i8 < lv.nEsc ? (1u << lv.p) + iL : i8;
  return tab[idx];
}

// Validates a weight set the way HUF_readDTableX1_wksp does (huf_decompress.c:140-149).
ZRA_DEV bool huf_check_weights(const u8* weights, u32 count) {
  u32 w1 = 0;
  for (u32 i = 0; i < count; i++) w1 += weights[i] == 1;
  return w1 >= 2 && !(w1 & 1);
}

// Fills the 1<<log entry single-symbol table: ascending weight, then ascending symbol.
ZRA_DEV bool huf_build_table(HufSym* table, const u8* weights, u32 count, u32 log) {
  u32 rank[16];
  for (u32 r = 0; r < 16; r++) rank[r] = 0;
  for (u32 i = 0; i < count; i++) rank[weights[i]]++;
  if (rank[1] < 2 || (rank[1] & 1)) return false;
  u32 start[16], nxt = 0;
  for (u32 r = 1; r <= log; r++) {
    start[r] = nxt;
    nxt += rank[r] << (r - 1);
  }
  for (u32 s = 0; s < count; s++) {
    u32 wt = weights[s];
    if (!wt) continue;
    u32 span = (1u << wt) >> 1;
    HufSym e = (HufSym)(s | ((log + 1 - wt) << 8));
    u32 b = start[wt];
    for (u32 u = 0; u < span; u++) table[b + u] = e;
    start[wt] = b + span;
  }
  return true;
}

}  // namespace zrab
