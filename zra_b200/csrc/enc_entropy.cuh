// enc_entropy.cuh — entropy-coding building blocks of the frame encoder (FSE + Huffman, zstd format).
//
// Thread-serial routines: one GPU thread prepares the tables of one block; the parallelism is
// across frames (and across the four Huffman streams of a block). Everything written here is the
// exact inverse of what entropy.cuh / decode_core.cuh read, so GPU-written frames decode with
// this library, the reference decoder and stock zstd alike. Reference behaviour mirrored
// (zstd/ = submodule/zstd/lib; the BYTES may differ from the reference encoder's, the format not):
//   forward bit writer ...... zstd/common/bitstream.h:180-270      (BIT_addBits / flush / close)
//   table-log choice ........ zstd/compress/fse_compress.c:325-342 (FSE_optimalTableLog)
//   count normalisation ..... zstd/compress/fse_compress.c:435-494 (FSE_normalizeCount: own algorithm here)
//   NCount header ........... zstd/compress/fse_compress.c:192-298 (FSE_writeNCount)
//   encode table ............ zstd/compress/fse_compress.c:66-169  (FSE_buildCTable_wksp)
//   symbol encode ........... zstd/common/fse.h:483-527            (FSE_initCState2 / encodeSymbol / flush)
//   Huffman lengths ......... zstd/compress/huf_compress.c:215-410 (HUF_buildCTable / setMaxHeight: own algorithm)
//   Huffman table header .... zstd/compress/huf_compress.c:114-147 (HUF_writeCTable)
#pragma once
#include "zfmt.cuh"

namespace zrab {

// ------------------------------------------------------------------ forward bit writer
// Bits are appended LSB-first; the stream is later read backwards, so the LAST bits written are the
// first ones decoded. close() appends the 1-bit end mark.
struct BitWriter {
  u8* out;
  u32 pos;   // bytes produced (keeps counting past `cap`, nothing is stored there)
  u32 cap;   // bytes that may be stored; a result > cap tells the caller the stream did not fit
  u64 acc;
  u32 nbits;

  ZRA_DEV void init(u8* dst, u32 capacity = 0xFFFFFFFFu) { out = dst; pos = 0; cap = capacity; acc = 0; nbits = 0; }
  ZRA_DEV void put(u8 b) {
    if (pos < cap) out[pos] = b;
    pos++;
  }
  ZRA_DEV void add(u32 value, u32 n) {  // n <= 32, value < 2^n
    acc |= (u64)value << nbits;
    nbits += n;
    while (nbits >= 8) {
      put((u8)acc);
      acc >>= 8;
      nbits -= 8;
    }
  }
  ZRA_DEV u32 close() {
    add(1, 1);
    if (nbits) { put((u8)acc); acc = 0; nbits = 0; }
    return pos;
  }
  // Plain byte-aligned finish without an end mark (NCount headers).
  ZRA_DEV u32 finish() {
    if (nbits) { put((u8)acc); acc = 0; nbits = 0; }
    return pos;
  }
};

// ------------------------------------------------------------------ FSE encode tables
struct FseSymTT {
  i32 deltaFindState;
  u32 deltaNbBits;
};

// Encode table of one alphabet: symbolTT[maxSymbol+1] + stateTable[1 << log].
struct FseCTable {
  FseSymTT* tt;    // per symbol
  u16* state;      // per table cell
  u32 log;
};

ZRA_DEV u32 fse_optimal_log(u32 maxLog, u32 total, u32 maxSymbol) {
  u32 maxBitsSrc = highbit32(total - 1) - 2;
  u32 log = maxLog;
  u32 minBitsSrc = highbit32(total) + 1, minBitsSym = highbit32(maxSymbol) + 2;
  u32 minBits = minBitsSrc < minBitsSym ? minBitsSrc : minBitsSym;
  if (maxBitsSrc < log) log = maxBitsSrc;
  if (minBits > log) log = minBits;
  if (log < 5) log = 5;
  if (log > 12) log = 12;
  return log;
}

// Normalises count[0..maxSymbol] (sum = total, at least two symbols present) to sum 1<<log with
// every present symbol >= 1. Largest-remainder style; the correction lands on the biggest symbols.
ZRA_DEV void fse_normalize(int16_t* norm, const u32* count, u32 maxSymbol, u32 total, u32 log) {
  const u32 size = 1u << log;
  u32 sum = 0, largest = 0;
  for (u32 s = 0; s <= maxSymbol; s++) {
    u32 c = count[s];
    if (!c) { norm[s] = 0; continue; }
    u64 scaled = ((u64)c << log) * 2 / total;  // fixed point with one fractional bit
    u32 p = (u32)((scaled + 1) >> 1);          // round to nearest
    if (p == 0) p = 1;
    norm[s] = (int16_t)p;
    sum += p;
    if (c > count[largest] || !count[largest]) largest = s;
  }
  while (sum != size) {
    if (sum < size) {
      norm[largest] = (int16_t)(norm[largest] + (size - sum));
      sum = size;
    } else {
      // take from the symbol with the largest normalised count that can give
      u32 best = largest;
      for (u32 s = 0; s <= maxSymbol; s++)
        if (norm[s] > norm[best]) best = s;
      u32 give = sum - size;
      u32 can = (u32)norm[best] - 1;
      if (can == 0) break;  // cannot happen while size >= number of symbols
      if (give > can / 2 + 1) give = can / 2 + 1;
      norm[best] = (int16_t)(norm[best] - give);
      sum -= give;
    }
  }
}

// Writes the NCount description of norm[0..maxSymbol]. Returns bytes written.
ZRA_DEV u32 fse_write_ncount(u8* dst, const int16_t* norm, u32 maxSymbol, u32 log) {
  BitWriter bw;
  bw.init(dst);
  const u32 alphabet = maxSymbol + 1;
  bw.add(log - 5, 4);
  i32 remaining = (i32)(1u << log) + 1;
  i32 threshold = (i32)(1u << log);
  u32 nbBits = log + 1;
  u32 sym = 0;
  bool previous0 = false;
  while (sym < alphabet && remaining > 1) {
    if (previous0) {
      u32 start = sym;
      while (sym < alphabet && !norm[sym]) sym++;
      if (sym == alphabet) break;
      while (sym >= start + 24) { start += 24; bw.add(0xFFFFu, 16); }
      while (sym >= start + 3) { start += 3; bw.add(3, 2); }
      bw.add(sym - start, 2);
    }
    i32 count = norm[sym++];
    i32 max = (2 * threshold - 1) - remaining;
    remaining -= count < 0 ? -count : count;
    count++;
    if (count >= threshold) count += max;
    bw.add((u32)count, nbBits - (count < max ? 1u : 0u));
    previous0 = (count == 1);
    while (remaining < threshold) { nbBits--; threshold >>= 1; }
  }
  return bw.finish();
}

// Builds the encode table from normalised counts (cells with count -1 are treated as 1).
// `cells` is scratch of 1<<log bytes.
ZRA_DEV void fse_build_ctable(FseCTable& ct, const int16_t* norm, u32 maxSymbol, u32 log, u8* cells) {
  const u32 size = 1u << log, mask = size - 1;
  u32 cumul[64];
  ct.log = log;
  u32 high = size - 1;
  cumul[0] = 0;
  for (u32 s = 0; s <= maxSymbol; s++) {
    if (norm[s] == -1) { cumul[s + 1] = cumul[s] + 1; cells[high--] = (u8)s; }
    else cumul[s + 1] = cumul[s] + (u32)norm[s];
  }
  u32 step = (size >> 1) + (size >> 3) + 3, pos = 0;
  for (u32 s = 0; s <= maxSymbol; s++) {
    for (i32 i = 0; i < norm[s]; i++) {
      cells[pos] = (u8)s;
      do { pos = (pos + step) & mask; } while (pos > high);
    }
  }
  for (u32 u = 0; u < size; u++) {
    u32 s = cells[u];
    ct.state[cumul[s]++] = (u16)(size + u);
  }
  u32 total = 0;
  for (u32 s = 0; s <= maxSymbol; s++) {
    i32 c = norm[s];
    if (c == 0) {
      ct.tt[s].deltaNbBits = ((log + 1) << 16) - size;
      ct.tt[s].deltaFindState = 0;
    } else if (c == -1 || c == 1) {
      ct.tt[s].deltaNbBits = (log << 16) - size;
      ct.tt[s].deltaFindState = (i32)total - 1;
      total++;
    } else {
      u32 maxBitsOut = log - highbit32((u32)c - 1);
      u32 minStatePlus = (u32)c << maxBitsOut;
      ct.tt[s].deltaNbBits = (maxBitsOut << 16) - minStatePlus;
      ct.tt[s].deltaFindState = (i32)total - c;
      total += (u32)c;
    }
  }
}

// Table for an RLE-coded alphabet: zero bits per symbol.
ZRA_DEV void fse_build_ctable_rle(FseCTable& ct, u32 symbol) {
  ct.log = 0;
  ct.state[0] = 0;
  ct.tt[symbol].deltaNbBits = 0;
  ct.tt[symbol].deltaFindState = 0;
}

ZRA_DEV u32 fse_init_state(const FseCTable& ct, u32 symbol) {
  if (ct.log == 0) return 0;
  const FseSymTT t = ct.tt[symbol];
  u32 nbBitsOut = (t.deltaNbBits + (1u << 15)) >> 16;
  u32 value = (nbBitsOut << 16) - t.deltaNbBits;
  return ct.state[(i32)(value >> nbBitsOut) + t.deltaFindState];
}
ZRA_DEV u32 fse_encode(const FseCTable& ct, BitWriter& bw, u32 state, u32 symbol) {
  if (ct.log == 0) return 0;
  const FseSymTT t = ct.tt[symbol];
  u32 nbBitsOut = (state + t.deltaNbBits) >> 16;
  bw.add(state & ((1u << nbBitsOut) - 1u), nbBitsOut);
  return ct.state[(i32)(state >> nbBitsOut) + t.deltaFindState];
}
ZRA_DEV void fse_flush_state(const FseCTable& ct, BitWriter& bw, u32 state) {
  bw.add(state & ((1u << ct.log) - 1u), ct.log);
}

// ------------------------------------------------------------------ Huffman
struct HufCode {
  u16 code;
  u8 len;
  u8 pad;
};

// Length-limited Huffman code lengths for count[0..maxSymbol] (at least two symbols present).
// Package-free construction: plain Huffman tree from a sorted symbol list, lengths clamped to
// maxBits, Kraft sum repaired by lengthening the cheapest symbols, slack handed back to the most
// frequent ones. Returns the longest length in use.
ZRA_DEV u32 huf_build_lengths(u8* len, const u32* count, u32 maxSymbol, u32 maxBits) {
  u16 order[256];
  u32 n = 0;
  for (u32 s = 0; s <= maxSymbol; s++) {
    len[s] = 0;
    if (count[s]) order[n++] = (u16)s;
  }
  // insertion sort by ascending count
  for (u32 i = 1; i < n; i++) {
    u16 v = order[i];
    u32 c = count[v];
    u32 j = i;
    while (j > 0 && count[order[j - 1]] > c) { order[j] = order[j - 1]; j--; }
    order[j] = v;
  }
  // two-queue merge: leaves [0,n) sorted, internal nodes [n, 2n-1) are produced in ascending order
  u32 weight[512];
  u16 parent[512];
  for (u32 i = 0; i < n; i++) weight[i] = count[order[i]];
  u32 leaf = 0, inner = n, next = n;
  while (next < 2 * n - 1) {
    u32 pick[2];
    for (u32 k = 0; k < 2; k++) {
      bool useLeaf = leaf < n && (inner >= next || weight[leaf] <= weight[inner]);
      pick[k] = useLeaf ? leaf++ : inner++;
    }
    weight[next] = weight[pick[0]] + weight[pick[1]];
    parent[pick[0]] = (u16)next;
    parent[pick[1]] = (u16)next;
    next++;
  }
  // depths from the root (the last node) down, clamped to maxBits as they propagate; every node
  // (leaf or internal) that had to be clamped is one unit of overflow — the classic deflate scheme
  u8 depth[512];
  u32 perLen[16];
  for (u32 b = 0; b < 16; b++) perLen[b] = 0;
  u32 overflow = 0;
  depth[2 * n - 2] = 0;
  for (i32 i = (i32)(2 * n - 3); i >= 0; i--) {
    u32 d = (u32)depth[parent[i]] + 1;
    if (d > maxBits) { d = maxBits; overflow++; }
    depth[i] = (u8)d;
    if ((u32)i < n) perLen[d]++;
  }
  // restore the Kraft equality: move one leaf down from the deepest level that still has room and
  // hang one of the overflowed leaves next to it (each move settles two units of overflow)
  while ((i32)overflow > 0) {
    u32 bits = maxBits - 1;
    while (perLen[bits] == 0) bits--;
    perLen[bits]--;
    perLen[bits + 1] += 2;
    perLen[maxBits]--;
    overflow -= 2;
  }
  // longest codes go to the least frequent symbols (order[] is ascending by count)
  u32 idx = 0, longest = 0;
  for (u32 bits = maxBits; bits >= 1; bits--) {
    for (u32 k = 0; k < perLen[bits]; k++) {
      len[order[idx++]] = (u8)bits;
      if (bits > longest) longest = bits;
    }
  }
  return longest;
}

// Canonical codes exactly as the decoder lays its table out: weight ascending (longest codes first),
// symbols ascending inside a weight; code = first table cell >> (tableLog - len).
ZRA_DEV void huf_assign_codes(HufCode* codes, const u8* len, u32 maxSymbol, u32 tableLog) {
  u32 rank[16];
  for (u32 r = 0; r < 16; r++) rank[r] = 0;
  for (u32 s = 0; s <= maxSymbol; s++) if (len[s]) rank[tableLog + 1 - len[s]]++;
  u32 start[16], nxt = 0;
  for (u32 w = 1; w <= tableLog; w++) { start[w] = nxt; nxt += rank[w] << (w - 1); }
  for (u32 s = 0; s <= maxSymbol; s++) {
    codes[s].pad = 0;
    if (!len[s]) { codes[s].code = 0; codes[s].len = 0; continue; }
    u32 w = tableLog + 1 - len[s];
    codes[s].code = (u16)(start[w] >> (tableLog - len[s]));
    codes[s].len = len[s];
    start[w] += 1u << (w - 1);
  }
}

// Writes the tree description (weights of symbols 0..maxSymbol-1; the last one is implied).
// Tries FSE-compressed weights, falls back to 4-bit direct weights. Returns bytes written, 0 when the
// tree cannot be described (more than 128 weights and incompressible) — the caller then stores
// the literals raw. `scratch` needs 1 KiB.
ZRA_DEV u32 huf_write_table(u8* dst, const u8* len, u32 maxSymbol, u32 tableLog, u8* scratch) {
  u8 weights[256];
  u32 wcount[16];
  for (u32 r = 0; r < 16; r++) wcount[r] = 0;
  const u32 n = maxSymbol;  // number of explicit weights
  u32 maxW = 0;
  for (u32 s = 0; s < n; s++) {
    u32 w = len[s] ? tableLog + 1 - len[s] : 0;
    weights[s] = (u8)w;
    wcount[w]++;
    if (w > maxW) maxW = w;
  }
  // ---- FSE-compressed weights (needs > 1 weights, not all equal, not all distinct)
  u32 mostCommon = 0;
  for (u32 w = 0; w <= maxW; w++) if (wcount[w] > mostCommon) mostCommon = wcount[w];
  if (n > 1 && mostCommon != n && mostCommon > 1) {
    u32 log = fse_optimal_log(6, n, maxW);
    if (log > 6) log = 6;
    int16_t norm[16];
    fse_normalize(norm, wcount, maxW, n, log);
    u32 h = fse_write_ncount(dst + 1, norm, maxW, log);
    FseSymTT tt[16];
    u16* st = reinterpret_cast<u16*>(scratch);        // 64 entries
    u8* cells = scratch + 128;                        // 64 bytes
    FseCTable ct;
    ct.tt = tt;
    ct.state = st;
    fse_build_ctable(ct, norm, maxW, log, cells);
    BitWriter bw;
    bw.init(dst + 1 + h);
    // two interleaved states: symbol i belongs to state (i & 1); encode from the last symbol back
    u32 s1 = 0, s2 = 0;
    bool have1 = false, have2 = false;
    for (i32 i = (i32)n - 1; i >= 0; i--) {
      if (i & 1) {
        if (!have2) { s2 = fse_init_state(ct, weights[i]); have2 = true; }
        else s2 = fse_encode(ct, bw, s2, weights[i]);
      } else {
        if (!have1) { s1 = fse_init_state(ct, weights[i]); have1 = true; }
        else s1 = fse_encode(ct, bw, s1, weights[i]);
      }
    }
    fse_flush_state(ct, bw, s2);
    fse_flush_state(ct, bw, s1);
    u32 body = bw.close();
    u32 total = h + body;
    if (total < 128 && total < (n + 1) / 2 + 0) {
      dst[0] = (u8)total;
      return total + 1;
    }
  }
  // ---- direct 4-bit weights
  if (n > 128) return 0;
  dst[0] = (u8)(127 + n);
  for (u32 s = 0; s < n; s += 2) {
    u32 a = weights[s], b = (s + 1 < n) ? weights[s + 1] : 0;
    dst[1 + s / 2] = (u8)((a << 4) | b);
  }
  return 1 + (n + 1) / 2;
}

// Encodes src[0..n) as one Huffman stream (symbols are emitted last-to-first so that the backward
// reader regenerates them in order). Returns bytes written.
ZRA_DEV u32 huf_encode_stream(u8* dst, u32 cap, const u8* src, u32 n, const HufCode* codes) {
  BitWriter bw;
  bw.init(dst, cap);
  for (u32 i = n; i-- > 0;) {
    HufCode c = codes[src[i]];
    bw.add(c.code, c.len);
  }
  return bw.close();
}

}  // namespace zrab
