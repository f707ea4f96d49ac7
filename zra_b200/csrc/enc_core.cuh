// enc_core.cuh — per-thread stages of the frame-parallel zstd ENCODER (levels 1-3: fast / dfast).
//
// A "round" encodes block r (<= 128 KiB) of every frame in flight, as a chain of small kernels
// whose per-thread bodies live here (thread-serial, so tests/host_sim can run them on the CPU):
//   enc_match     (1 thread / frame)   greedy hash match finder -> packed (ll, ml, offsetValue) records
//   enc_literals  (1 thread / frame)   gathers the literal bytes, byte histogram
//   enc_plan      (1 thread / frame)   literal mode + Huffman code, sequence modes + FSE tables,
//                                      section headers
//   enc_huf       (1 thread / stream)  Huffman streams
//   enc_seq       (1 thread / frame)   FSE sequence bitstream
//   enc_assemble  (1 thread / frame)   block header + sections (or a raw block), frame header,
//                                      checksum slot, frame size
// What the reference does on this path (zstd/ = submodule/zstd/lib/compress), mirrored in spirit —
// same greedy heuristics and parameters, not the same bytes:
//   parameters .......... zstd_compress.c:4125-4252, 1031-1062   (cParams table + adjust)
//   fast matcher ........ zstd_fast.c:46-183
//   dfast matcher ....... zstd_double_fast.c:50-316
//   literals ............ zstd_compress_literals.c:70-158
//   sequences ........... zstd_compress.c:1967-2152, zstd_compress_sequences.c:141-359
//   block / frame ....... zstd_compress.c:2418-2478, 2586-2704, 3179-3216
#pragma once
#include "decode_core.cuh"   // seq_pack / seq_ll / seq_ml / seq_off, FrameDesc-style helpers
#include "enc_entropy.cuh"
#include "xxh64.cuh"

namespace zrab {

struct EncParams {
  u32 hashLogS;   // short-hash table log (mls bytes)
  u32 hashLogL;   // long-hash (8 bytes) table log; 0 = single-table "fast" strategy
  u32 mls;        // bytes hashed by the short hash (4..7)
  u32 step;       // extra skip step (fast strategy "acceleration")
  u32 checksum;
  u32 windowLogMax;
};

// Level -> parameters for a frame of `frameLen` bytes (the reference's table rows, see file header).
ZRA_DEV EncParams enc_params(int level, u32 frameLen, bool checksum) {
  if (level == 0) level = 3;
  u32 accel = 0;
  if (level < 0) { accel = (u32)(-level); level = 0; }
  if (level > 4) level = 4;  // greedy/lazy/bt* parsers are not implemented: strongest available
  // rows: {windowLog, chainLog, hashLog, minMatch, dfast?}
  const u32 cls = frameLen <= (16u << 10) ? 3 : (frameLen <= (128u << 10) ? 2 : (frameLen <= (256u << 10) ? 1 : 0));
  u32 W, C, H, L, D;
  if (cls == 3)      { W = 14; C = level == 0 ? 12 : 14; H = level == 0 ? 13 : 15; L = level <= 1 ? 5 : 4; D = level >= 3; }
  else if (cls == 2) { W = 17; u32 c[5] = {12, 12, 13, 15, 16}, h[5] = {12, 13, 15, 16, 17}, l[5] = {5, 6, 5, 5, 5}; C = c[level]; H = h[level]; L = l[level]; D = level >= 3; }
  else if (cls == 1) { W = 18; u32 c[5] = {12, 13, 14, 16, 16}, h[5] = {13, 14, 14, 16, 17}, l[5] = {5, 6, 5, 4, 4}; C = c[level]; H = h[level]; L = l[level]; D = level >= 2; }
  else               { u32 w[5] = {19, 19, 20, 21, 21}, c[5] = {12, 13, 15, 16, 18}, h[5] = {13, 14, 16, 17, 18}, l[5] = {6, 7, 6, 5, 5}; W = w[level]; C = c[level]; H = h[level]; L = l[level]; D = level >= 3; }
  u32 srcLog = frameLen > 64 ? highbit32(frameLen - 1) + 1 : 6;
  if (W > srcLog) W = srcLog;
  if (H > W + 1) H = W + 1;
  if (D && C > W) C = W;
  if (W < 10) W = 10;
  EncParams p;
  p.hashLogS = D ? C : H;
  p.hashLogL = D ? H : 0;
  p.mls = L < 4 ? 4 : (L > 7 ? 7 : L);
  p.step = accel;
  p.checksum = checksum ? 1 : 0;
  // The matchers look for candidates anywhere in the frame (a frame is its own window), so the DECLARED window must
  // cover the whole frame (RFC 8878 3.1.1.1.2: no offset beyond Window_Size) even where the level's row has a smaller
  // one — frames above 512 KiB / 1 MiB / 2 MiB at levels <= 1 / 2 / >= 3. One-shot decoders never looked, streaming
  // decoders (zstd -d, ZSTD_decompressStream) size their buffers by it.
  p.windowLogMax = W > srcLog ? W : (srcLog < 10 ? 10 : srcLog);
  return p;
}

// Mutable per-frame encoder state (HBM), one per frame in flight.
struct EncCtx {
  u32 status;
  u32 srcLen;        // frame length
  u32 blkPos;        // offset of the current block in the frame
  u32 blkLen;        // its length
  u32 lastBlock;
  u32 rep[3];        // repeat-offset history as the DECODER will see it
  u32 outPos;        // bytes of the frame's output slot written so far
  // ---- current block
  u32 nbSeq, litSize;
  u32 litMode;       // 0 raw, 1 rle, 2 huffman
  u32 hufLog, hufHeaderSize, nStreams;
  u32 hufStreamSize[4];
  u32 seqModes[3];   // LL, OF, ML: 0 predefined, 1 rle, 2 compressed
  u32 seqLog[3];     // table logs of the three encode tables
  u32 repSave[3];    // history at the start of the block (restored if the block is stored raw)
  u32 seqHeaderSize; // nbSeq field + modes byte + table descriptions
  u32 seqStreamSize;
  u32 windowLog;
  u32 blkActive;     // 0 when this round has no block for the frame (short last frame)
};

// Device scratch of one frame for the encoder (pointers into HBM regions).
struct EncScratch {
  u32* tabS;         // 1 << hashLogS entries (position + 1, 0 = empty)
  u32* tabL;         // 1 << hashLogL entries
  u64* seqs;         // packed sequence records of the current block
  u8* lit;           // literal bytes of the current block
  u32* hist;         // 256 literal counts
  HufCode* hcodes;   // 256
  u8* hufOut;        // 4 regions of hufStride bytes
  u32 hufStride;
  u8* hdr;           // literal-section tree description + sequence section header (512 bytes)
  FseSymTT* tt;      // 36 + 32 + 53 entries (LL, OF, ML)
  u16* states;       // 512 + 256 + 512
  u8* seqOut;        // sequence bitstream
  u32 seqOutCap;
  u8* cells;         // 1 KiB of spread / weight-coding scratch
  u32* cnt;          // LL[36] | OF[32] | ML[53] code counts when the matcher already made them, else null
};

// ------------------------------------------------------------------ unaligned reads of the input
// `base` is the 4-byte aligned start of the whole input buffer; `off` an absolute byte offset.
ZRA_DEV u32 in32(const u8* base, u64 off) {
  const u32* w = reinterpret_cast<const u32*>(base) + (off >> 2);
  u32 sh = (u32)(off & 3) * 8;
  u32 lo = w[0];
  if (!sh) return lo;
  return fsh_rc(lo, w[1], sh);
}
ZRA_DEV u64 in64(const u8* base, u64 off) { return (u64)in32(base, off) | ((u64)in32(base, off + 4) << 32); }

ZRA_DEV u32 hash_short(const u8* base, u64 off, u32 mls, u32 hlog) {
  if (mls <= 4) return (in32(base, off) * 2654435761u) >> (32 - hlog);
  u64 v = in64(base, off) << (64 - 8 * mls);
  return (u32)((v * 0x9E3779B97F4A7C15ull) >> (64 - hlog));
}
ZRA_DEV u32 hash_long(const u8* base, u64 off, u32 hlog) {
  return (u32)((in64(base, off) * 0xCF1BBCDCB7A56463ull) >> (64 - hlog));
}

// Number of equal bytes of [a, ...) and [b, ...), a < limit (frame-relative offsets, fbase = frame start).
ZRA_DEV u32 match_len(const u8* base, u64 fbase, u32 a, u32 b, u32 limit) {
  u32 n = 0;
  while (a + n + 4 <= limit) {
    u32 x = in32(base, fbase + a + n) ^ in32(base, fbase + b + n);
    if (x) {
#if defined(__CUDA_ARCH__)
      return n + ((u32)__ffs((int)x) - 1) / 8;
#else
      return n + (u32)__builtin_ctz(x) / 8;
#endif
    }
    n += 4;
  }
  while (a + n < limit && base[fbase + a + n] == base[fbase + b + n]) n++;
  return n;
}

// ------------------------------------------------------------------ sequence emission
// Converts (litLength, matchLength, real offset) to the wire offsetValue with the decoder's
// repeat-offset rules and keeps the history in step with what the decoder will compute.
ZRA_DEV u64 emit_sequence(EncCtx& c, u32 ll, u32 ml, u32 offset) {
  // Branch-free (this runs in the serial selection loop of the frame-cooperative matcher). With ll > 0 the wire values
  // 1/2/3 mean rep0/rep1/rep2; with ll == 0 they mean rep1/rep2/rep0-1 (zstd_decompress_block.c:871-888).
  const u32 r0 = c.rep[0], r1 = c.rep[1], r2 = c.rep[2];
  const bool ll0 = ll == 0;
  const u32 c1 = ll0 ? r1 : r0, c2 = ll0 ? r2 : r1, c3 = ll0 ? r0 - 1u : r2;
  const bool m1 = offset == c1;
  const bool m2 = !m1 && offset == c2;
  const bool m3 = !m1 && !m2 && offset == c3 && !(ll0 && r0 <= 1u);
  const u32 idx = m1 ? 1u : (m2 ? 2u : (m3 ? 3u : 0u));
  const u32 value = idx ? idx : offset + 3u;
  // history: rep0 itself (idx 1, ll > 0) changes nothing; rep1 (idx + ll0 == 2) swaps the first two; the rest shifts
  const bool same = m1 && !ll0;
  const bool keep2 = idx != 0 && idx + (ll0 ? 1u : 0u) <= 2u;
  c.rep[2] = keep2 ? r2 : r1;
  c.rep[1] = same ? r1 : r0;
  c.rep[0] = offset;
  return seq_pack(ll, ml, value);
}

// ------------------------------------------------------------------ enc_match (1 thread / frame)
// Greedy parse of the current block. Positions are frame-relative; the window is the whole frame.
ZRA_DEV void enc_match(const u8* base, u64 fbase, const EncParams& p, EncCtx& c, const EncScratch& s) {
  c.nbSeq = 0;
  c.litSize = 0;
  c.repSave[0] = c.rep[0]; c.repSave[1] = c.rep[1]; c.repSave[2] = c.rep[2];
  const u32 bstart = c.blkPos, bend = c.blkPos + c.blkLen;
  if (c.blkLen < 16) { c.litSize = c.blkLen; return; }
  const u32 ilimit = bend - 8;  // last position whose 8 bytes can be hashed
  const bool dfast = p.hashLogL != 0;
  u32 ip = bstart, anchor = bstart;
  if (ip == 0) ip = 1;  // position 0 has nothing before it
  u32 off1 = c.rep[0], off2 = c.rep[1];
  // repeat offsets that would reach before the frame start are not usable
  if (off1 > ip) off1 = 0;
  if (off2 > ip) off2 = 0;
  u32 nseq = 0;
  while (ip < ilimit) {
    u32 mpos = 0, mlen = 0, moff = 0;  // match start, length, offset
    const u32 cur = ip;
    const u32 hs = hash_short(base, fbase + ip, p.mls, p.hashLogS);
    const u32 candS = s.tabS[hs];
    s.tabS[hs] = cur + 1;
    u32 candL = 0;
    if (dfast) {
      const u32 hl = hash_long(base, fbase + ip, p.hashLogL);
      candL = s.tabL[hl];
      s.tabL[hl] = cur + 1;
    }
    // 1. repeat offset one byte ahead (cheapest to code)
    if (off1 && in32(base, fbase + ip + 1 - off1) == in32(base, fbase + ip + 1)) {
      mpos = ip + 1;
      moff = off1;
      mlen = 4 + match_len(base, fbase, mpos + 4, mpos + 4 - moff, bend);
    } else if (dfast && candL && in64(base, fbase + candL - 1) == in64(base, fbase + ip)) {
      // 2. long (8-byte) candidate
      mpos = ip;
      moff = ip - (candL - 1);
      mlen = 8 + match_len(base, fbase, ip + 8, candL - 1 + 8, bend);
    } else if (candS && in32(base, fbase + candS - 1) == in32(base, fbase + ip)) {
      // 3. short candidate; with two tables, first see whether ip+1 has a long match
      mpos = ip;
      moff = ip - (candS - 1);
      bool taken = false;
      if (dfast) {
        const u32 hl1 = hash_long(base, fbase + ip + 1, p.hashLogL);
        const u32 c1 = s.tabL[hl1];
        s.tabL[hl1] = cur + 2;
        if (c1 && in64(base, fbase + c1 - 1) == in64(base, fbase + ip + 1)) {
          mpos = ip + 1;
          moff = mpos - (c1 - 1);
          mlen = 8 + match_len(base, fbase, mpos + 8, c1 - 1 + 8, bend);
          taken = true;
        }
      }
      if (!taken) mlen = 4 + match_len(base, fbase, ip + 4, candS - 1 + 4, bend);
    } else {
      ip += ((ip - anchor) >> (dfast ? 8 : 7)) + 1 + p.step;
      continue;
    }
    // extend backwards over the pending literals
    if (moff != off1 || mpos != ip + 1) {
      while (mpos > anchor && mpos - moff > 0 && base[fbase + mpos - 1] == base[fbase + mpos - moff - 1]) { mpos--; mlen++; }
    }
    // emit
    {
      u32 ll = mpos - anchor;
      s.seqs[nseq++] = emit_sequence(c, ll, mlen, moff);
      if (moff != off1) { off2 = off1; off1 = moff; }
    }
    ip = mpos + mlen;
    anchor = ip;
    if (ip <= ilimit) {
      // complementary insertions so that the positions just skipped can be found later
      const u32 a = cur + 2, b = ip - 2;
      if (a + 8 <= bend) {
        s.tabS[hash_short(base, fbase + a, p.mls, p.hashLogS)] = a + 1;
        if (dfast) s.tabL[hash_long(base, fbase + a, p.hashLogL)] = a + 1;
      }
      if (b > cur) {
        s.tabS[hash_short(base, fbase + (dfast ? ip - 1 : b), p.mls, p.hashLogS)] = (dfast ? ip - 1 : b) + 1;
        if (dfast) s.tabL[hash_long(base, fbase + b, p.hashLogL)] = b + 1;
      }
      // immediate repeat of the second offset (zero literals)
      while (ip <= ilimit && off2 && in32(base, fbase + ip) == in32(base, fbase + ip - off2)) {
        u32 rl = 4 + match_len(base, fbase, ip + 4, ip + 4 - off2, bend);
        u32 t = off2; off2 = off1; off1 = t;
        s.tabS[hash_short(base, fbase + ip, p.mls, p.hashLogS)] = ip + 1;
        if (dfast) s.tabL[hash_long(base, fbase + ip, p.hashLogL)] = ip + 1;
        s.seqs[nseq++] = emit_sequence(c, 0, rl, off1);
        ip += rl;
        anchor = ip;
      }
    }
  }
  c.nbSeq = nseq;
  c.litSize = 0;  // filled by enc_literals
  (void)bstart;
}

// ------------------------------------------------------------------ enc_literals (1 thread / frame)
ZRA_DEV void enc_literals(const u8* base, u64 fbase, EncCtx& c, const EncScratch& s) {
  for (u32 k = 0; k < 256; k++) s.hist[k] = 0;
  u32 pos = c.blkPos, n = 0;
  for (u32 i = 0; i < c.nbSeq; i++) {
    u32 ll = seq_ll(s.seqs[i]), ml = seq_ml(s.seqs[i]);
    for (u32 k = 0; k < ll; k++) {
      u8 b = base[fbase + pos + k];
      s.lit[n++] = b;
      s.hist[b]++;
    }
    pos += ll + ml;
  }
  const u32 bend = c.blkPos + c.blkLen;
  while (pos < bend) {
    u8 b = base[fbase + pos++];
    s.lit[n++] = b;
    s.hist[b]++;
  }
  c.litSize = n;
}

// ------------------------------------------------------------------ sequence codes
ZRA_DEV u32 ll_code(u32 ll) {
  if (ll < 16) return ll;
  if (ll < 64) {
    const u8 t[48] = {16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 20, 20, 21, 21, 21, 21, 22, 22, 22, 22, 22, 22, 22, 22,
                      23, 23, 23, 23, 23, 23, 23, 23, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24};
    return t[ll - 16];
  }
  return highbit32(ll) + 19;
}
ZRA_DEV u32 ml_code(u32 mlBase) {  // mlBase = matchLength - 3
  if (mlBase < 32) return mlBase;
  if (mlBase < 128) {
    // codes 32..42 cover 32..127 (see kMLBase / kMLBits)
    u32 m = mlBase + 3;
    u32 code = 32;
    while (code < 42 && m >= kMLBase[code + 1]) code++;
    return code;
  }
  return highbit32(mlBase) + 36;
}
ZRA_DEV u32 of_code(u32 offsetValue) { return highbit32(offsetValue); }

// ------------------------------------------------------------------ enc_plan (1 thread / frame)
// Decides literal and sequence coding modes, builds all tables, writes the two section headers
// into s.hdr: [0, hufHeaderSize) tree description, [256, 256 + seqHeaderSize) sequence header.
ZRA_DEV void enc_plan(EncCtx& c, const EncScratch& s) {
  // ---------------- literals
  const u32 n = c.litSize;
  c.litMode = 0;
  c.nStreams = 0;
  c.hufHeaderSize = 0;
  if (n > 63) {
    u32 maxSym = 0, largest = 0;
    for (u32 k = 0; k < 256; k++) {
      if (s.hist[k]) { maxSym = k; if (s.hist[k] > largest) largest = s.hist[k]; }
    }
    if (largest == n) {
      c.litMode = 1;
    } else if (largest > (n >> 7) + 4) {
      u32 maxBits = 11;
      u32 bySrc = highbit32(n - 1) - 1;
      if (bySrc < maxBits) maxBits = bySrc;
      u32 minBits = highbit32(maxSym) + 2;
      if (highbit32(n) + 1 < minBits) minBits = highbit32(n) + 1;
      if (maxBits < minBits) maxBits = minBits;
      if (maxBits < 5) maxBits = 5;
      if (maxBits > 11) maxBits = 11;
      u8 len[256];
      u32 log = huf_build_lengths(len, s.hist, maxSym, maxBits);
      huf_assign_codes(s.hcodes, len, maxSym, log);
      for (u32 k = maxSym + 1; k < 256; k++) { s.hcodes[k].code = 0; s.hcodes[k].len = 0; s.hcodes[k].pad = 0; }
      u32 h = huf_write_table(s.hdr, len, maxSym, log, s.cells);
      if (h) {
        u64 bits = 0;
        for (u32 k = 0; k <= maxSym; k++) bits += (u64)s.hist[k] * len[k];
        u32 streams = n < 256 ? 1 : 4;
        u32 est = h + (u32)((bits + 7) >> 3) + (streams == 4 ? 6 + 3 : 1);
        u32 minGain = (n >> 6) + 2;
        if (est + minGain < n) {
          c.litMode = 2;
          c.hufLog = log;
          c.hufHeaderSize = h;
          c.nStreams = streams;
        }
      }
    }
  }
  // ---------------- sequences
  const u32 nbSeq = c.nbSeq;
  u8* sh = s.hdr + 256;
  u32 hp = 0;
  if (nbSeq < 128) sh[hp++] = (u8)nbSeq;
  else if (nbSeq < kLongNbSeq) { sh[hp++] = (u8)((nbSeq >> 8) + 0x80); sh[hp++] = (u8)nbSeq; }
  else { sh[hp++] = 0xFF; sh[hp++] = (u8)(nbSeq - kLongNbSeq); sh[hp++] = (u8)((nbSeq - kLongNbSeq) >> 8); }
  c.seqHeaderSize = hp;
  if (!nbSeq) return;
  u32 cntLL[36], cntOF[32], cntML[53];
  for (u32 k = 0; k < 36; k++) cntLL[k] = 0;
  for (u32 k = 0; k < 32; k++) cntOF[k] = 0;
  for (u32 k = 0; k < 53; k++) cntML[k] = 0;
  if (s.cnt) {
    for (u32 k = 0; k < 36; k++) cntLL[k] = s.cnt[k];
    for (u32 k = 0; k < 32; k++) cntOF[k] = s.cnt[36 + k];
    for (u32 k = 0; k < 53; k++) cntML[k] = s.cnt[68 + k];
  } else {
    for (u32 i = 0; i < nbSeq; i++) {
      u64 q = s.seqs[i];
      cntLL[ll_code(seq_ll(q))]++;
      cntOF[of_code(seq_off(q))]++;
      cntML[ml_code(seq_ml(q) - 3)]++;
    }
  }
  const u32 modesPos = hp++;
  u32 modes = 0;
  for (u32 t = 0; t < 3; t++) {  // wire order: LL, OF, ML
    u32* cnt = t == 0 ? cntLL : (t == 1 ? cntOF : cntML);
    const u32 alphabet = t == 0 ? 36 : (t == 1 ? 32 : 53);
    const u32 maxLog = t == 1 ? kOFFSELog : kLLFSELog;
    const u32 defLog = t == 1 ? kOFDefLog : kLLDefLog;
    const int16_t* defNorm = t == 0 ? kLLDefNorm : (t == 1 ? kOFDefNorm : kMLDefNorm);
    const u32 defMax = t == 0 ? kMaxLL : (t == 1 ? kDefaultMaxOF : kMaxML);
    FseCTable ct;
    ct.tt = s.tt + (t == 0 ? 0 : (t == 1 ? 36 : 68));
    ct.state = s.states + (t == 0 ? 0 : (t == 1 ? 512 : 768));
    u32 maxSym = 0, most = 0, present = 0;
    for (u32 k = 0; k < alphabet; k++) if (cnt[k]) { maxSym = k; present++; if (cnt[k] > most) most = cnt[k]; }
    u32 mode;
    const bool defaultOk = maxSym <= defMax;
    if (present == 1) mode = 1;
    else if (defaultOk && (nbSeq < ((1u << defLog) * 9u >> 3) || most < (nbSeq >> (defLog - 1)))) mode = 0;
    else mode = 2;
    if (mode == 1) {
      sh[hp++] = (u8)maxSym;
      fse_build_ctable_rle(ct, maxSym);
    } else if (mode == 0) {
      fse_build_ctable(ct, defNorm, defMax, defLog, s.cells);
    } else {
      // the last sequence's symbols ride in the initial states: leave them out of the statistics
      u64 lastq = s.seqs[nbSeq - 1];
      u32 lastSym = t == 0 ? ll_code(seq_ll(lastq)) : (t == 1 ? of_code(seq_off(lastq)) : ml_code(seq_ml(lastq) - 3));
      u32 total = nbSeq;
      if (cnt[lastSym] > 1) { cnt[lastSym]--; total--; }
      u32 log = fse_optimal_log(maxLog, total, maxSym);
      if (log > maxLog) log = maxLog;
      int16_t norm[64];
      fse_normalize(norm, cnt, maxSym, total, log);
      hp += fse_write_ncount(sh + hp, norm, maxSym, log);
      fse_build_ctable(ct, norm, maxSym, log, s.cells);
    }
    c.seqModes[t] = mode;
    c.seqLog[t] = ct.log;
    modes |= mode << (6 - 2 * t);
  }
  sh[modesPos] = (u8)modes;
  c.seqHeaderSize = hp;
}

// ------------------------------------------------------------------ enc_huf (1 thread / stream)
ZRA_DEV void enc_huf(EncCtx& c, const EncScratch& s, u32 stream) {
  if (c.litMode != 2 || stream >= c.nStreams) return;
  u32 seg = c.nStreams == 4 ? (c.litSize + 3) / 4 : c.litSize;
  u32 beg = stream * seg;
  u32 n = (c.nStreams == 4 && stream == 3) ? c.litSize - 3 * seg : seg;
  c.hufStreamSize[stream] = huf_encode_stream(s.hufOut + (u64)stream * s.hufStride, s.hufStride, s.lit + beg, n, s.hcodes);
}

// ------------------------------------------------------------------ enc_seq (1 thread / frame)
ZRA_DEV void enc_seq(EncCtx& c, const EncScratch& s) {
  c.seqStreamSize = 0;
  const u32 nbSeq = c.nbSeq;
  if (!nbSeq) return;
  FseCTable ctLL, ctOF, ctML;
  ctLL.tt = s.tt; ctLL.state = s.states;
  ctOF.tt = s.tt + 36; ctOF.state = s.states + 512;
  ctML.tt = s.tt + 68; ctML.state = s.states + 768;
  ctLL.log = c.seqLog[0];
  ctOF.log = c.seqLog[1];
  ctML.log = c.seqLog[2];
  BitWriter bw;
  bw.init(s.seqOut, s.seqOutCap);
  u64 q = s.seqs[nbSeq - 1];
  u32 ll = seq_ll(q), ml = seq_ml(q) - 3, ov = seq_off(q);
  u32 lc = ll_code(ll), mc = ml_code(ml), oc = of_code(ov);
  u32 stML = fse_init_state(ctML, mc);
  u32 stOF = fse_init_state(ctOF, oc);
  u32 stLL = fse_init_state(ctLL, lc);
  bw.add(ll - kLLBase[lc], kLLBits[lc]);
  bw.add(ml + 3 - kMLBase[mc], kMLBits[mc]);
  bw.add(ov - (1u << oc), oc);
  for (u32 i = nbSeq - 1; i-- > 0;) {
    q = s.seqs[i];
    ll = seq_ll(q); ml = seq_ml(q) - 3; ov = seq_off(q);
    lc = ll_code(ll); mc = ml_code(ml); oc = of_code(ov);
    stOF = fse_encode(ctOF, bw, stOF, oc);
    stML = fse_encode(ctML, bw, stML, mc);
    stLL = fse_encode(ctLL, bw, stLL, lc);
    bw.add(ll - kLLBase[lc], kLLBits[lc]);
    bw.add(ml + 3 - kMLBase[mc], kMLBits[mc]);
    bw.add(ov - (1u << oc), oc);
  }
  fse_flush_state(ctML, bw, stML);
  fse_flush_state(ctOF, bw, stOF);
  fse_flush_state(ctLL, bw, stLL);
  c.seqStreamSize = bw.close();
}

// ------------------------------------------------------------------ enc_assemble (1 thread / frame)
// Appends the current block to the frame's output slot; on the first block also the frame header,
// after the last block the checksum. `out` = start of the frame's slot.
ZRA_DEV void enc_assemble(const u8* base, u64 fbase, const EncParams& p, EncCtx& c, const EncScratch& s, u8* out) {
  u32 op = c.outPos;
  if (c.blkPos == 0) {
    out[0] = 0x28; out[1] = 0xB5; out[2] = 0x2F; out[3] = 0xFD;
    out[4] = (u8)(p.checksum << 2);
    out[5] = (u8)((c.windowLog - 10) << 3);
    op = 6;
  }
  const u32 n = c.litSize;
  // literals section size
  u32 litSection;
  u32 cLit = 0;
  if (c.litMode == 2) {
    cLit = c.hufHeaderSize + (c.nStreams == 4 ? 6 : 0);
    for (u32 k = 0; k < c.nStreams; k++) cLit += c.hufStreamSize[k];
    u32 lh = 3 + (n >= 1024) + (n >= 16384);
    litSection = lh + cLit;
    // the estimate in enc_plan can be off by a few bytes: re-check the gain with the real size
    bool spilled = false;
    for (u32 k = 0; k < c.nStreams; k++) spilled |= c.hufStreamSize[k] > s.hufStride;
    if (spilled || litSection + ((n >> 6) + 2) >= n + 3 || cLit >= (1u << 18) || (c.nStreams == 4 && (c.hufStreamSize[0] > 65535 || c.hufStreamSize[1] > 65535 || c.hufStreamSize[2] > 65535))) {
      c.litMode = 0;
    }
  }
  if (c.litMode == 0) litSection = (n < 32 ? 1 : (n < 4096 ? 2 : 3)) + n;
  else if (c.litMode == 1) litSection = (n < 32 ? 1 : (n < 4096 ? 2 : 3)) + 1;
  const u32 seqSection = c.seqHeaderSize + c.seqStreamSize;
  const u32 cSize = litSection + seqSection;
  const u32 minGain = (c.blkLen >> 6) + 2;
  bool raw = c.blkLen < 7 || cSize + minGain >= c.blkLen || cSize >= kBlockSizeMax || c.seqStreamSize > s.seqOutCap;
  // RLE block, the reference's rule (zstd_compress.c:2453-2464): not the first block, "maybe RLE" sequence store, all bytes equal
  bool rleBlock = false;
  if (c.blkPos != 0 && c.nbSeq < 4 && c.litSize < 10 && c.blkLen > 0) {
    rleBlock = true;
    for (u32 k = 1; k < c.blkLen; k++) rleBlock &= base[fbase + c.blkPos + k] == base[fbase + c.blkPos];
  }
  if (rleBlock) raw = false;
  const u32 bsize = (raw || rleBlock) ? c.blkLen : cSize;
  const u32 bh = (c.lastBlock ? 1u : 0u) | ((rleBlock ? 1u : (raw ? 0u : 2u)) << 1) | (bsize << 3);
  out[op++] = (u8)bh; out[op++] = (u8)(bh >> 8); out[op++] = (u8)(bh >> 16);
  if (rleBlock) {
    out[op++] = base[fbase + c.blkPos];
    c.rep[0] = c.repSave[0]; c.rep[1] = c.repSave[1]; c.rep[2] = c.repSave[2];
  } else if (raw) {
    for (u32 k = 0; k < c.blkLen; k++) out[op++] = base[fbase + c.blkPos + k];
    // a raw block carries no sequences: undo the repeat-offset updates the matcher made for it
    c.rep[0] = c.repSave[0]; c.rep[1] = c.repSave[1]; c.rep[2] = c.repSave[2];
  } else {
    // literals section header
    if (c.litMode == 2) {
      u32 single = c.nStreams == 1;
      if (n < 1024) {
        u32 v = 2u | ((single ? 0u : 1u) << 2) | (n << 4) | (cLit << 14);
        out[op++] = (u8)v; out[op++] = (u8)(v >> 8); out[op++] = (u8)(v >> 16);
      } else if (n < 16384) {
        u32 v = 2u | (2u << 2) | (n << 4) | (cLit << 18);
        out[op++] = (u8)v; out[op++] = (u8)(v >> 8); out[op++] = (u8)(v >> 16); out[op++] = (u8)(v >> 24);
      } else {
        u64 v = 2u | (3u << 2) | ((u64)n << 4) | ((u64)cLit << 22);
        for (u32 k = 0; k < 5; k++) out[op++] = (u8)(v >> (8 * k));
      }
      for (u32 k = 0; k < c.hufHeaderSize; k++) out[op++] = s.hdr[k];
      if (c.nStreams == 4) {
        for (u32 k = 0; k < 3; k++) { out[op++] = (u8)c.hufStreamSize[k]; out[op++] = (u8)(c.hufStreamSize[k] >> 8); }
      }
      for (u32 st = 0; st < c.nStreams; st++) {
        const u8* hs = s.hufOut + (u64)st * s.hufStride;
        for (u32 k = 0; k < c.hufStreamSize[st]; k++) out[op++] = hs[k];
      }
    } else {
      u32 type = c.litMode;  // 0 raw, 1 rle
      if (n < 32) out[op++] = (u8)(type | (n << 3));
      else if (n < 4096) { u32 v = type | (1u << 2) | (n << 4); out[op++] = (u8)v; out[op++] = (u8)(v >> 8); }
      else { u32 v = type | (3u << 2) | (n << 4); out[op++] = (u8)v; out[op++] = (u8)(v >> 8); out[op++] = (u8)(v >> 16); }
      if (type == 1) out[op++] = s.lit[0];
      else for (u32 k = 0; k < n; k++) out[op++] = s.lit[k];
    }
    // sequences section
    for (u32 k = 0; k < c.seqHeaderSize; k++) out[op++] = s.hdr[256 + k];
    for (u32 k = 0; k < c.seqStreamSize; k++) out[op++] = s.seqOut[k];
  }
  c.outPos = op;
}

// XXH64 of the whole frame by one thread (host logic tests; the kernel uses the quad-lane form).
ZRA_DEV u32 enc_checksum_serial(const u8* base, u64 fbase, u32 len) {
  u64 acc[4];
  for (u32 q = 0; q < 4; q++) {
    acc[q] = xxh_init_acc(q);
    for (u32 k = 0; k < (len >> 5); k++) acc[q] = xxh_round(acc[q], ld64(base + fbase + 32 * (u64)k + 8 * q));
  }
  u64 h;
  if (len >= 32) {
    h = xxh_rotl(acc[0], 1) + xxh_rotl(acc[1], 7) + xxh_rotl(acc[2], 12) + xxh_rotl(acc[3], 18);
    for (u32 q = 0; q < 4; q++) h = xxh_merge(h, acc[q]);
  } else {
    h = kXP5;
  }
  return (u32)xxh_finish(h, len, base + fbase + (len & ~31u), len & 31u);
}

}  // namespace zrab
