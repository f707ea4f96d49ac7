// decode_core.cuh — per-thread stages of the frame-parallel zstd decoder.
//
// The decoder is a pipeline of small kernels; a "round" decodes block r of every frame in flight:
//   block_setup  (1 thread / frame)   frame+block headers, literal section header, Huffman table,
//                                     sequence section header, three FSE tables
//   huf_stream   (1 thread / stream)  Huffman literals -> literal scratch (4 streams per block)
//   seq_decode   (1 thread / frame)   FSE sequence decode -> packed (ll, ml, offset) records,
//                                     full validation, repcode history, block output size
//   seq_execute  (1 warp / frame)     literal + match copies into the output (decode_kernels.cu)
//   frame_finish (4 threads / frame)  XXH64 content checksum, size checks (decode_kernels.cu)
// The serial entropy work is thread-per-frame on purpose: a warp advances 32 independent frames
// in lock-step, so every issue slot does useful work and the parallelism is the archive's frame
// count. Reference behaviour being reproduced (zstd/ = submodule/zstd/lib):
//   frame header ........ zstd/decompress/zstd_decompress.c:244-318
//   block loop .......... zstd/decompress/zstd_decompress.c:609-692, zstd_decompress_block.c:56-70
//   literals section .... zstd/decompress/zstd_decompress_block.c:79-235
//   sequence headers .... zstd/decompress/zstd_decompress_block.c:433-550
//   sequence decode ..... zstd/decompress/zstd_decompress_block.c:795-948
//   Huffman streams ..... zstd/decompress/huf_decompress.c:239-354
#pragma once
#include "entropy.cuh"

namespace zrab {

// One frame to decode: where its bytes are and where its output goes.
struct FrameDesc {
  u64 srcOff;   // byte offset of the zstd frame in the source buffer
  u64 dstOff;   // byte offset of its output in the destination buffer
  u32 srcLen;   // compressed size of the frame
  u32 dstCap;   // output capacity
  u32 exact;    // 1: the frame must regenerate exactly dstCap bytes (ZRA seek-table geometry)
  u32 pad;
};

enum BlockType : u32 { BT_RAW = 0, BT_RLE = 1, BT_COMPRESSED = 2, BT_NONE = 3 };
enum LitMode : u32 { LIT_RAW = 0, LIT_RLE = 1, LIT_HUF = 2 };
enum FrameFlags : u32 { FF_DONE = 1, FF_CHECKSUM = 2, FF_HUF_VALID = 4, FF_FSE_VALID = 8, FF_STARTED = 16, FF_FINISHED = 32 };

// Mutable per-frame decoder state, resident in HBM across rounds.
struct FrameCtx {
  u32 status;   // ZErr; sticky
  u32 flags;
  u32 srcPos;   // offset (within the frame) of the next block header
  u32 dstPos;   // bytes regenerated so far
  u32 rep[3];
  u32 hufLog;
  u64 fcs;      // frame content size, ~0 when absent
  // ---- current block (rewritten by block_setup every round)
  u32 blkType;
  u32 blkSrc;   // offset of the block content within the frame
  u32 blkSize;  // raw: bytes, rle: regenerated size, compressed: compressed size
  u32 blkDst;   // offset (within the frame's output) where this block starts
  u32 blkOut;   // regenerated size of this block
  u32 litMode, litSrc, litSize;
  u32 nStreams;
  u32 strOff[4], strLen[4];
  u32 nbSeq, seqOff, seqLen;
  u32 llLog, ofLog, mlLog;
  u32 hufCount;  // number of Huffman weights (FrameTables::hufWeights), the implied last one included
};

// Table scratch of one frame (HBM). The sequence tables use the compact 2-byte format.
constexpr u32 kHufGlobalCap = 4096 + 256;  // two-level table of any legal code (log <= 12): L1 256 + L2 < 4096
struct FrameTables {
  CSym ll[512];
  CSym ml[512];
  CSym of[256];
  u8 hufWeights[256];        // weights of the current Huffman tree (kept across blocks for treeless literals)
  HufSym huf[kHufGlobalCap]; // only used by frames whose two-level table does not fit its shared-memory slot
};

ZRA_DEV void frame_fail(FrameCtx& c, u32 code) {
  if (!c.status) c.status = code;
  c.blkType = BT_NONE;
  c.flags |= FF_DONE;
}

// ---------------------------------------------------------------- frame header
ZRA_DEV bool parse_frame_header(const u8* f, const FrameDesc& d, FrameCtx& c) {
  c.status = 0; c.flags = FF_STARTED; c.srcPos = 0; c.dstPos = 0;
  c.rep[0] = 1; c.rep[1] = 4; c.rep[2] = 8;
  c.hufLog = 0; c.fcs = ~0ull; c.blkType = BT_NONE; c.blkOut = 0;
  if (d.srcLen < 6 + 3) { frame_fail(c, ZE_SRC_WRONG); return false; }
  if (ld32(f) != kZstdMagic) { frame_fail(c, ZE_PREFIX_UNKNOWN); return false; }
  u32 fhd = f[4];
  u32 didCode = fhd & 3, single = (fhd >> 5) & 1, fcsId = fhd >> 6;
  u32 didSz = didCode == 3 ? 4 : didCode;
  u32 fcsSz = fcsId == 0 ? 0 : (1u << fcsId);
  u32 hsz = 5 + !single + didSz + fcsSz + (u32)(single && !fcsId);
  if (d.srcLen < hsz + 3) { frame_fail(c, ZE_SRC_WRONG); return false; }
  if (fhd & 8) { frame_fail(c, ZE_FRAMEPARAM_UNSUPPORTED); return false; }
  u32 pos = 5;
  if (!single) {
    u32 wl = (f[pos++] >> 3) + 10;
    if (wl > 31) { frame_fail(c, ZE_WINDOW_TOO_LARGE); return false; }
  }
  u32 dictID = 0;
  if (didCode == 1) dictID = f[pos]; else if (didCode == 2) dictID = ld16(f + pos); else if (didCode == 3) dictID = ld32(f + pos);
  pos += didSz;
  if (fcsId == 0) { if (single) c.fcs = f[pos]; }
  else if (fcsId == 1) c.fcs = ld16(f + pos) + 256;
  else if (fcsId == 2) c.fcs = ld32(f + pos);
  else c.fcs = ld64(f + pos);
  if (dictID) { frame_fail(c, ZE_DICT_WRONG); return false; }
  if (fhd & 4) c.flags |= FF_CHECKSUM;
  c.srcPos = hsz;
  return true;
}

// ---------------------------------------------------------------- sequence table descriptor
// Returns bytes consumed, or 0xFFFFFFFF on error.
ZRA_DEV u32 setup_seq_table(CSym* table, u32* logOut, u32 type, u32 kind, const u8* src, u32 len, bool repeatOk) {
  const u32 maxSym = kind == SEQ_LL ? kMaxLL : (kind == SEQ_ML ? kMaxML : kMaxOF);
  const u32 maxLog = kind == SEQ_OF ? kOFFSELog : kLLFSELog;
  if (type == 1) {  // RLE
    if (!len || src[0] > maxSym) return 0xFFFFFFFFu;
    fse_build_compact_rle(table, src[0]);
    *logOut = 0;
    return 1;
  }
  if (type == 0) {  // predefined
    if (kind == SEQ_LL) { fse_build_compact(table, kLLDefNorm, kMaxLL, kLLDefLog); *logOut = kLLDefLog; }
    else if (kind == SEQ_ML) { fse_build_compact(table, kMLDefNorm, kMaxML, kMLDefLog); *logOut = kMLDefLog; }
    else { fse_build_compact(table, kOFDefNorm, kDefaultMaxOF, kOFDefLog); *logOut = kOFDefLog; }
    return 0;
  }
  if (type == 3) return repeatOk ? 0 : 0xFFFFFFFFu;  // keep the previous block's table
  int16_t norm[64];
  u32 ms = maxSym, log, err = 0;
  u32 h = fse_read_ncount(src, len, norm, &ms, &log, &err);
  if (!h || log > maxLog) return 0xFFFFFFFFu;
  fse_build_compact(table, norm, ms, log);
  *logOut = log;
  return h;
}

// ---------------------------------------------------------------- block_setup (1 thread / frame)
// block_setup in three steps, so that a kernel can build ONE sequence table at a time in a small shared-memory
// staging area and copy it out (warp-cooperatively) before the next: head (headers, literals section, sequence
// section header), table x 3 in stream order (LL, OF, ML), tail. `SetupCursor` carries the parse position between
// the steps; live == false means the block needs no (more) table work — finished, failed, raw, RLE or sequence-free.
struct SetupCursor {
  u32 ip, iend;   // offsets within the frame of the next unread byte / the end of the block content
  u32 modes, nbSeq;
  bool live;
};

ZRA_DEV SetupCursor block_setup_head(const u8* srcBase, const FrameDesc& d, FrameCtx& c, FrameTables& t, bool firstRound) {
  SetupCursor cur;
  cur.ip = cur.iend = cur.modes = cur.nbSeq = 0;
  cur.live = false;
  const u8* f = srcBase + d.srcOff;
  if (firstRound) { if (!parse_frame_header(f, d, c)) return cur; }
  c.blkType = BT_NONE;
  if (c.status || (c.flags & FF_DONE)) return cur;
  u32 tail = (c.flags & FF_CHECKSUM) ? 4u : 0u;
  (void)tail;
  if (c.srcPos + 3 > d.srcLen) { frame_fail(c, ZE_SRC_WRONG); return cur; }
  u32 bh = ld24(f + c.srcPos);
  u32 last = bh & 1, type = (bh >> 1) & 3, bsz = bh >> 3;
  if (type == 3) { frame_fail(c, ZE_CORRUPTION); return cur; }
  u32 csz = (type == BT_RLE) ? 1 : bsz;
  u32 content = c.srcPos + 3;
  if (content + csz > d.srcLen) { frame_fail(c, ZE_SRC_WRONG); return cur; }
  c.blkSrc = content;
  c.blkSize = bsz;
  c.blkDst = c.dstPos;
  c.srcPos = content + csz;
  if (last) c.flags |= FF_DONE;
  if (type != BT_COMPRESSED) {
    if (bsz > d.dstCap - c.dstPos) { frame_fail(c, ZE_DST_TOO_SMALL); return cur; }
    c.blkType = type;
    c.blkOut = bsz;
    c.dstPos += bsz;
    return cur;
  }
  // ---- compressed block
  if (csz >= kBlockSizeMax) { frame_fail(c, ZE_SRC_WRONG); return cur; }
  if (csz < 3) { frame_fail(c, ZE_CORRUPTION); return cur; }
  const u8* b = f + content;
  u32 used;
  c.nStreams = 0;
  {
    u32 ltype = b[0] & 3, fmt = (b[0] >> 2) & 3;
    if (ltype >= 2) {
      u32 lh, cs, litSize;
      bool single = false;
      if (ltype == 3 && !(c.flags & FF_HUF_VALID)) { frame_fail(c, ZE_DICT_CORRUPTED); return cur; }
      if (csz < 5) { frame_fail(c, ZE_CORRUPTION); return cur; }
      u32 w = ld32(b);
      if (fmt <= 1) { single = !fmt; lh = 3; litSize = (w >> 4) & 0x3FF; cs = (w >> 14) & 0x3FF; }
      else if (fmt == 2) { lh = 4; litSize = (w >> 4) & 0x3FFF; cs = w >> 18; }
      else { lh = 5; litSize = (w >> 4) & 0x3FFFF; cs = (w >> 22) + ((u32)b[4] << 10); }
      if (litSize > kBlockSizeMax || cs + lh > csz) { frame_fail(c, ZE_CORRUPTION); return cur; }
      // every literal ends up in the output: a literals section larger than what is left of the frame cannot decode
      // (zstd fails it with dstSize_tooSmall once the bytes are copied out, zstd_decompress_block.c:1049-1051), and
      // the per-frame literal scratch is sized by the frame, so it must be refused BEFORE the Huffman stage runs
      if (litSize > d.dstCap - c.dstPos) { frame_fail(c, ZE_DST_TOO_SMALL); return cur; }
      u32 hoff = content + lh, hlen = cs;
      if (ltype == 2) {
        u8 weights[256];
        u32 count, log;
        u32 h = huf_read_weights(srcBase, d.srcOff + hoff, hlen, weights, &count, &log);
        if (!h || !huf_check_weights(weights, count)) { frame_fail(c, ZE_CORRUPTION); return cur; }
        // the decode table itself is built by the Huffman stage, straight into shared memory
        for (u32 i = 0; i < count; i++) t.hufWeights[i] = weights[i];
        c.hufCount = count;
        c.hufLog = log;
        c.flags |= FF_HUF_VALID;
        hoff += h; hlen -= h;
      }
      if (single) {
        c.nStreams = 1;
        c.strOff[0] = hoff; c.strLen[0] = hlen;
      } else {
        if (hlen < 10) { frame_fail(c, ZE_CORRUPTION); return cur; }
        const u8* j = f + hoff;
        u32 l1 = ld16(j), l2 = ld16(j + 2), l3 = ld16(j + 4);
        if (l1 + l2 + l3 + 6 > hlen) { frame_fail(c, ZE_CORRUPTION); return cur; }
        u32 seg = (litSize + 3) / 4;
        if (seg * 3 > litSize) { frame_fail(c, ZE_CORRUPTION); return cur; }
        c.nStreams = 4;
        c.strOff[0] = hoff + 6; c.strLen[0] = l1;
        c.strOff[1] = c.strOff[0] + l1; c.strLen[1] = l2;
        c.strOff[2] = c.strOff[1] + l2; c.strLen[2] = l3;
        c.strOff[3] = c.strOff[2] + l3; c.strLen[3] = hlen - 6 - l1 - l2 - l3;
      }
      c.litMode = LIT_HUF; c.litSrc = 0; c.litSize = litSize;
      used = lh + cs;
    } else {
      u32 lh, litSize;
      if (fmt == 0 || fmt == 2) { lh = 1; litSize = b[0] >> 3; }
      else if (fmt == 1) { lh = 2; litSize = ld16(b) >> 4; }
      else { lh = 3; litSize = ld24(b) >> 4; }
      if (ltype == 0) {
        if (lh + litSize > csz) { frame_fail(c, ZE_CORRUPTION); return cur; }
        c.litMode = LIT_RAW; c.litSrc = content + lh; c.litSize = litSize;
        used = lh + litSize;
      } else {
        if (lh + 1 > csz || litSize > kBlockSizeMax) { frame_fail(c, ZE_CORRUPTION); return cur; }
        c.litMode = LIT_RLE; c.litSrc = b[lh]; c.litSize = litSize;
        used = lh + 1;
      }
    }
  }
  // ---- sequences section header
  const u8* ip = b + used;
  const u8* iend = b + csz;
  if (ip >= iend) { frame_fail(c, ZE_SRC_WRONG); return cur; }
  u32 nbSeq = *ip++;
  if (!nbSeq) {
    if (ip != iend) { frame_fail(c, ZE_SRC_WRONG); return cur; }
  } else {
    if (nbSeq > 0x7F) {
      if (nbSeq == 0xFF) {
        if (ip + 2 > iend) { frame_fail(c, ZE_SRC_WRONG); return cur; }
        nbSeq = ld16(ip) + kLongNbSeq; ip += 2;
      } else {
        if (ip >= iend) { frame_fail(c, ZE_SRC_WRONG); return cur; }
        nbSeq = ((nbSeq - 0x80) << 8) + *ip++;
      }
    }
    if (ip + 1 > iend) { frame_fail(c, ZE_SRC_WRONG); return cur; }
    cur.modes = *ip++;
    cur.live = true;
  }
  cur.nbSeq = nbSeq;
  cur.ip = (u32)(ip - f);
  cur.iend = (u32)(iend - f);
  return cur;
}

// One table (kind in stream order: SEQ_LL, SEQ_OF, SEQ_ML) into `dst` (the frame's FrameTables array, or a staging
// copy of it). Returns true when `dst` was (re)built; a "repeat" table keeps the previous block's copy.
ZRA_DEV bool block_setup_table(const u8* srcBase, const FrameDesc& d, FrameCtx& c, SetupCursor& cur, u32 kind, CSym* dst) {
  if (!cur.live) return false;
  const u8* f = srcBase + d.srcOff;
  const u32 type = kind == SEQ_LL ? cur.modes >> 6 : (kind == SEQ_OF ? (cur.modes >> 4) & 3 : (cur.modes >> 2) & 3);
  u32* logOut = kind == SEQ_LL ? &c.llLog : (kind == SEQ_OF ? &c.ofLog : &c.mlLog);
  const bool rep = (c.flags & FF_FSE_VALID) != 0;
  const u32 h = setup_seq_table(dst, logOut, type, kind, f + cur.ip, cur.iend - cur.ip, rep);
  if (h == 0xFFFFFFFFu) { frame_fail(c, ZE_CORRUPTION); cur.live = false; return false; }
  cur.ip += h;
  return type != 3;
}

ZRA_DEV void block_setup_tail(const FrameDesc& d, FrameCtx& c, SetupCursor& cur) {
  if (c.status || c.blkType == BT_RAW || c.blkType == BT_RLE) return;
  if (cur.nbSeq) {
    if (!cur.live) return;  // failed in a table
    c.flags |= FF_FSE_VALID;
    if (cur.ip >= cur.iend) { frame_fail(c, ZE_CORRUPTION); return; }  // the bitstream needs at least its end mark
  } else if (cur.iend == 0) {
    return;  // the head stopped before the sequence section (frame done earlier, nothing to do)
  }
  c.nbSeq = cur.nbSeq;
  c.seqOff = cur.ip;
  c.seqLen = cur.iend - cur.ip;
  c.blkType = BT_COMPRESSED;
  c.blkOut = 0;
  if (!cur.nbSeq) {  // literals only: nothing for the sequence stage to do
    if (c.litSize > d.dstCap - c.blkDst) { frame_fail(c, ZE_DST_TOO_SMALL); return; }
    c.blkOut = c.litSize;
    c.dstPos = c.blkDst + c.litSize;
  }
}

// Thread-serial whole (host logic tests; frames handled outside the staging kernel).
ZRA_DEV void block_setup(const u8* srcBase, const FrameDesc& d, FrameCtx& c, FrameTables& t, bool firstRound) {
  SetupCursor cur = block_setup_head(srcBase, d, c, t, firstRound);
  block_setup_table(srcBase, d, c, cur, SEQ_LL, t.ll);
  block_setup_table(srcBase, d, c, cur, SEQ_OF, t.of);
  block_setup_table(srcBase, d, c, cur, SEQ_ML, t.ml);
  block_setup_tail(d, c, cur);
}

// ---------------------------------------------------------------- bit reader
// Backward bit reader of the sequence stream and of the Huffman streams, built for 32 unrelated streams advancing in
// lock-step in one warp. SIMT facts that shape it (measured, profiles/r01b): a data-dependent
// refill branch is taken by SOME lane at every step, so its body runs every step with two or
// three lanes active; and a per-lane register prefetch does not work, because the scoreboard is
// per warp — one lane's outstanding load stalls every other lane that touches the same register.
// So the reader carries NO window between steps, only a bit position:
//  * compressed bytes are staged in a small per-lane ring in shared memory (16 words, laid out
//    [word][lane] so that a warp access is always bank-conflict free);
//  * every step rebuilds a 64-bit window from three ring words at the current bit position
//    (3 LDS + 2 funnel shifts, unconditional, no branch);
//  * the ring is topped up warp-synchronously at "refill points" (every 2 steps): each lane
//    issues at most one 16-byte LDG for its next group and stores the group it issued at the
//    previous point — every lane issues and consumes at the same instruction, so the per-warp
//    scoreboard never couples unrelated lanes, and the load has two steps to land.
// Bit indices are absolute within the frame's "group space": group 0 is the 16-byte aligned
// group that holds the stream's first byte, word w = 4*group + i. Unread bits are [b0, p).
// Nothing is masked: an over-read consumes whatever lies below the stream (or stale ring words)
// and is caught by p != b0 at the end of the block; garbage states stay inside their tables.
constexpr u32 kRingWords = 16;

struct SeqReader {
  i32 p;                 // absolute index of the next unread bit, plus one
  i32 b0;                // absolute index of the stream's first bit (0..127)
  i32 loadedW;           // lowest word index present in the ring
  i32 reqW;              // lowest word index requested (== loadedW when no group is in flight)
  const u32* g0;         // address of group 0 (16-byte aligned)
  u32 f0, f1, f2, f3;    // the group in flight (words reqW .. reqW+3)
  u32* ring;             // this lane's ring: word w lives at ring[(w & 15) * stride]
  u32 stride;

  ZRA_DEV u32 ring_ld(i32 w) const { return ring[((u32)w & (kRingWords - 1)) * stride]; }
  ZRA_DEV void ring_st(i32 w, u32 v) { ring[((u32)w & (kRingWords - 1)) * stride] = v; }
  ZRA_DEV void fetch_group(i32 firstWord) {  // f <- words firstWord .. firstWord+3
#if defined(__CUDA_ARCH__)
    uint4 v = __ldg(reinterpret_cast<const uint4*>(g0 + firstWord));
    f0 = v.x; f1 = v.y; f2 = v.z; f3 = v.w;
#else
    f0 = g0[firstWord]; f1 = g0[firstWord + 1]; f2 = g0[firstWord + 2]; f3 = g0[firstWord + 3];
#endif
  }
  ZRA_DEV void land() {  // the group in flight goes into the ring
    if (reqW < loadedW) {
      ring_st(reqW, f0); ring_st(reqW + 1, f1); ring_st(reqW + 2, f2); ring_st(reqW + 3, f3);
      loadedW = reqW;
    }
  }
  // Refill point. Between two points a lane consumes at most 4 words (2 steps x 64 bits); one
  // group per point keeps at least 9 landed words ahead of the cursor.
  ZRA_DEV void refill_point() {
    land();
    const i32 k = (p - 1) >> 5;
    // safety net (never taken by streams whose steps stay within 64 bits): synchronous top-up
    while (loadedW > 0 && k - loadedW < 6) {
      reqW = loadedW - 4;
      fetch_group(reqW);
      land();
    }
    if (reqW > 0 && k - reqW + 5 <= (i32)kRingWords) {
      reqW -= 4;
      fetch_group(reqW);
#if defined(__CUDA_ARCH__)
      if (reqW >= 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(g0 + reqW - 8));
#endif
    }
  }
  // The stream is src[byteOff, byteOff+len); base must be 16-byte aligned and the buffer readable
  // to the end of the 16-byte group that holds the stream's last byte.
  ZRA_DEV bool init(const u8* base, u64 byteOff, u32 len, u32* ringMem, u32 ringStride) {
    ring = ringMem; stride = ringStride;
    if (len == 0) return false;
    u32 last = base[byteOff + len - 1];
    if (last == 0) return false;
    const u8* first = base + byteOff;
    const u8* grp = reinterpret_cast<const u8*>(reinterpret_cast<uintptr_t>(first) & ~(uintptr_t)15);
    g0 = reinterpret_cast<const u32*>(grp);
    b0 = (i32)(first - grp) * 8;
    p = b0 + (i32)((len - 1) * 8 + highbit32(last));
    // synchronous initial fill: the top groups, as many as the ring holds
    const i32 topGroup = (p - 1 >= 0 ? (p - 1) : 0) >> 7;
    i32 g = topGroup;
    for (u32 n = 0; n < kRingWords / 4 && g >= 0; n++, g--) {
      fetch_group(4 * g);
      ring_st(4 * g, f0); ring_st(4 * g + 1, f1); ring_st(4 * g + 2, f2); ring_st(4 * g + 3, f3);
    }
    loadedW = reqW = 4 * (g + 1);
    return true;
  }
  // 64-bit window at the cursor: bit 31 of hi is the next unread bit.
  ZRA_DEV void window(u32& hi, u32& lo) const {
    const i32 t = p - 1;
    const i32 k = t >> 5;
    const u32 s = ~(u32)t & 31u;
    const u32 w0 = ring_ld(k), w1 = ring_ld(k - 1), w2 = ring_ld(k - 2);
    hi = fsh_lc(w1, w0, s);
    lo = fsh_lc(w2, w1, s);
  }
};

// n <= 32 bits off the top of a 64-bit window (hi:lo)
ZRA_DEV u32 win_take(u32& hi, u32& lo, u32 n) {
  u32 v = fsh_lc(hi, 0, n);
  hi = fsh_lc(lo, hi, n);
  lo = fsh_lc(0, lo, n);
  return v;
}

// ---------------------------------------------------------------- Huffman literals
// Four symbols off a 64-bit window; returns them packed little-endian, *used = bits consumed.
// At most 4 x 12 bits, so one window serves the group.
ZRA_DEV u32 huf_group4(const HufSym* tab, const HufLevels& lv, u32 hi, u32 lo, u32* used) {
  u32 word = 0, total = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (u32 j = 0; j < 4; j++) {
    const u32 e = huf_lookup(tab, lv, hi);
    const u32 nb = e >> 8;
    word |= (e & 0xFFu) << (8 * j);
    hi = fsh_lc(lo, hi, nb);
    lo = fsh_lc(0, lo, nb);
    total += nb;
  }
  *used = total;
  return word;
}

// Thread-serial decode of stream `s` of the current block of one frame into its literal scratch,
// with the same reader, table and group structure as k_huf_decode (host logic tests). 0 or a ZErr.
ZRA_DEV u32 huf_stream(const u8* srcBase, const FrameDesc& d, const FrameCtx& c, const HufSym* tab, const HufLevels& lv, u8* lit,
                       u32 s) {
  u32 seg = c.nStreams == 4 ? (c.litSize + 3) / 4 : c.litSize;
  u32 n = (c.nStreams == 4 && s == 3) ? c.litSize - 3 * seg : seg;
  u32 ring[kRingWords];
  SeqReader br;
  if (!br.init(srcBase, d.srcOff + c.strOff[s], c.strLen[s], ring, 1)) return ZE_CORRUPTION;
  u8* out = lit + s * seg;
  u32 i = 0, it = 0;
  while (i < n) {
    if ((it++ & 1) == 0) br.refill_point();
    u32 hi, lo, used;
    br.window(hi, lo);
    if (n - i >= 4) {
      u32 w = huf_group4(tab, lv, hi, lo, &used);
      out[i] = (u8)w; out[i + 1] = (u8)(w >> 8); out[i + 2] = (u8)(w >> 16); out[i + 3] = (u8)(w >> 24);
      i += 4;
    } else {
      const u32 e = huf_lookup(tab, lv, hi);
      used = e >> 8;
      out[i++] = (u8)e;
    }
    br.p -= (i32)used;
  }
  return br.p == br.b0 ? ZE_OK : ZE_CORRUPTION;
}

// ---------------------------------------------------------------- seq_decode (1 thread / frame)
// Sequence records handed to seq_execute are CUMULATIVE, so the executor needs no prefix scans:
//   litEnd (18 bits) | outEnd (18 bits) << 18 | offset (28 bits) << 36
// litEnd = literals consumed by sequences 0..i of the block, outEnd = bytes regenerated by them
// (both <= 128 KiB = 2^17, hence 18 bits), offset = the resolved match distance (repcodes done).
ZRA_DEV u64 seq_pack(u32 ll, u32 ml, u32 off) { return (u64)ll | ((u64)ml << 18) | ((u64)off << 36); }
ZRA_DEV u32 seq_ll(u64 s) { return (u32)s & 0x3FFFFu; }
ZRA_DEV u32 seq_ml(u64 s) { return (u32)(s >> 18) & 0x3FFFFu; }
ZRA_DEV u32 seq_off(u64 s) { return (u32)(s >> 36); }
ZRA_DEV u32 rec_lit_end(u64 s) { return (u32)s & 0x3FFFFu; }
ZRA_DEV u32 rec_out_end(u64 s) { return (u32)(s >> 18) & 0x3FFFFu; }
ZRA_DEV u32 rec_off(u64 s) { return (u32)(s >> 36); }
constexpr u32 kMaxOffset = (1u << 28) - 1;

// Per-frame state of the sequence stage; lives in registers of the lane that owns the frame.
struct SeqState {
  SeqReader br;
  u32 sLL, sML, sOF;       // FSE states
  u32 rep0, rep1, rep2;    // repeat-offset history
  u32 litUsed, produced;   // literals consumed / bytes regenerated so far in this block
  u32 i, n;                // sequence cursor / count
  u32 llLog, ofLog, mlLog;
  u32 litSize, room, blkDst;
  u32 err;                 // first error seen (sticky), ZErr
};

// Starts the sequence stage of the current block of one frame. Returns 0 or a ZErr.
ZRA_DEV u32 seq_begin(const u8* srcBase, const FrameDesc& d, const FrameCtx& c, u32 seqCap, SeqState& s, u32* ringMem,
                      u32 ringStride) {
  if (c.nbSeq > seqCap) return ZE_CORRUPTION;
  if (!s.br.init(srcBase, d.srcOff + c.seqOff, c.seqLen, ringMem, ringStride)) return ZE_CORRUPTION;
  s.llLog = c.llLog; s.ofLog = c.ofLog; s.mlLog = c.mlLog;
  u32 hi, lo;
  s.br.window(hi, lo);
  s.sLL = win_take(hi, lo, s.llLog);
  s.sOF = win_take(hi, lo, s.ofLog);
  s.sML = win_take(hi, lo, s.mlLog);
  s.br.p -= (i32)(s.llLog + s.ofLog + s.mlLog);
  s.rep0 = c.rep[0]; s.rep1 = c.rep[1]; s.rep2 = c.rep[2];
  s.litUsed = 0; s.produced = 0; s.i = 0; s.n = c.nbSeq;
  s.litSize = c.litSize; s.blkDst = c.blkDst;
  // a block regenerates at most Block_Maximum_Size bytes (format doc :329-417); the cumulative record fields rely on it
  s.room = d.dstCap - c.blkDst < kBlockSizeMax ? d.dstCap - c.blkDst : kBlockSizeMax;
  s.err = 0;
  return ZE_OK;
}

ZRA_DEV u32 sel32(bool c, u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
  u32 r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.b32 %0, %1, %2, p;\n\t}" : "=r"(r) : "r"(a), "r"(b), "r"((u32)c));
  return r;
#else
  return c ? a : b;
#endif
}

// Decodes and validates ONE sequence, completely branch-free in the common case (lanes of a warp
// run this in lock-step on unrelated frames, so every data-dependent branch would serialise).
// tLL/tML/tOF are the compact tables (shared memory in the kernel), lutLL/lutML the packed
// baseline tables. Errors are sticky in s.err (the first one wins, in the reference's order of
// checks); decoding simply continues on garbage, which cannot leave the tables or the record
// array. Reference: ZSTD_decodeSequence, zstd/decompress/zstd_decompress_block.c:838-948.
ZRA_DEV u64 seq_step(const CSym* tLL, const CSym* tML, const CSym* tOF, const u32* lutLL, const u32* lutML, SeqState& s) {
  SeqReader& br = s.br;
  u32 hi, lo;
  br.window(hi, lo);
  const u32 eLL = tLL[s.sLL], eML = tML[s.sML], eOF = tOF[s.sOF];
  const u32 lutl = lutLL[eLL & 63u], lutm = lutML[eML & 63u];
  const u32 ofc = eOF & 63u;
  const u32 llBase = lutl & 0xFFFFFFu, llBits = lutl >> 24;
  const u32 mlBase = lutm & 0xFFFFFFu, mlBits = lutm >> 24;
  // next-state parameters: nbBits = log - highbit(ns), base = (ns << nbBits) - tableSize
  const u32 keep = (s.i + 1 < s.n) ? 0xFFFFFFFFu : 0u;  // the last sequence leaves the states alone
  const u32 nsLL = eLL >> 6, nsML = eML >> 6, nsOF = eOF >> 6;
  const u32 nbLL = (s.llLog - highbit32(nsLL)) & keep;
  const u32 nbML = (s.mlLog - highbit32(nsML)) & keep;
  const u32 nbOF = (s.ofLog - highbit32(nsOF)) & keep;
  // ---- bits, in stream order: offset extra | match extra | literal extra | LL, ML, OF state bits
  const u32 extra = ofc + mlBits + llBits;        // <= 31 + 16 + 16
  const u32 stateBits = nbLL + nbML + nbOF;       // <= 26
  const u32 ofx = win_take(hi, lo, ofc);
  const u32 mlx = win_take(hi, lo, mlBits);
  const u32 llx = win_take(hi, lo, llBits);
  if (extra + stateBits > 64) {  // cannot happen below 2^25-byte windows with sane lengths: second window for the states
    br.p -= (i32)extra;
    br.window(hi, lo);
    br.p += (i32)extra;
  }
  const u32 bLL = win_take(hi, lo, nbLL);
  const u32 bML = win_take(hi, lo, nbML);
  const u32 bOF = win_take(hi, lo, nbOF);
  br.p -= (i32)(extra + stateBits);
  s.sLL = ((nsLL << nbLL) - (1u << s.llLog) + bLL) & keep;
  s.sML = ((nsML << nbML) - (1u << s.mlLog) + bML) & keep;
  s.sOF = ((nsOF << nbOF) - (1u << s.ofLog) + bOF) & keep;
  // ---- offset: new value or one of the three repeat offsets (zstd_decompress_block.c:871-888)
  const u32 ll = llBase + llx, ml = mlBase + mlx;
  const u32 ll0 = (llBase == 0);
  const bool isRep = ofc <= 1;
  const u32 idx = ofc + ofx + ll0;                      // only meaningful when isRep: 0..3
  const u32 newOff = (1u << (ofc & 31u)) - 3u + ofx;    // only meaningful when !isRep
  // select chain, not branches: the lanes of a warp are unrelated frames, a taken branch here serialises them
  // (profiles/r02i_source_k_seq_decode.txt: 6.7 % of the instructions, 11.4 % of the stall samples on this line)
  u32 repv = sel32(idx == 0, s.rep0, sel32(idx == 1, s.rep1, sel32(idx == 2, s.rep2, s.rep0 - 1u)));
  repv += !repv;
  const u32 offset = isRep ? repv : newOff;
  if (!isRep || idx >= 2) s.rep2 = s.rep1;
  if (!isRep || idx >= 1) s.rep1 = s.rep0;
  s.rep0 = offset;
  // ---- validation: everything seq_execute will trust
  const u32 litEnd = s.litUsed + ll;
  const u32 outEnd = s.produced + ll + ml;
  u32 e = 0;
  if (offset > s.blkDst + s.produced + ll || offset > kMaxOffset) e = ZE_CORRUPTION;
  if (ll + ml > s.room - s.produced) e = ZE_DST_TOO_SMALL;  // produced <= room is an invariant while err == 0
  if (ll > s.litSize - s.litUsed) e = ZE_CORRUPTION;
  if (!s.err) s.err = e;
  if (!e) { s.litUsed = litEnd; s.produced = outEnd; }   // keeps the invariants (and the record fields) in range
  s.i++;
  return (u64)(litEnd & 0x3FFFFu) | ((u64)(outEnd & 0x3FFFFu) << 18) | ((u64)(offset & kMaxOffset) << 36);
}

// Finishes the block after its last sequence: stream exhaustion, trailing literals, write-back.
ZRA_DEV u32 seq_end(const SeqState& s, FrameCtx& c) {
  if (s.err) return s.err;
  if (s.br.p != s.br.b0) return ZE_CORRUPTION;
  u32 lastLits = s.litSize - s.litUsed;
  if (lastLits > s.room - s.produced) return ZE_DST_TOO_SMALL;
  c.rep[0] = s.rep0; c.rep[1] = s.rep1; c.rep[2] = s.rep2;
  c.blkOut = s.produced + lastLits;
  c.dstPos = s.blkDst + c.blkOut;
  return ZE_OK;
}

// Whole stage for one frame, thread-serial (host logic tests; the kernel interleaves 32 frames).
ZRA_DEV void seq_decode(const u8* srcBase, const FrameDesc& d, FrameCtx& c, const FrameTables& t, u64* seqs, u32 seqCap) {
  if (c.blkType != BT_COMPRESSED || c.status || !c.nbSeq) return;
  u32 lutLL[64], lutML[64], ring[kRingWords];
  for (u32 k = 0; k < 64; k++) { lutLL[k] = k < 36 ? ll_lut(k) : 0; lutML[k] = k < 53 ? ml_lut(k) : 0; }
  SeqState s;
  u32 err = seq_begin(srcBase, d, c, seqCap, s, ring, 1);
  while (!err && s.i < s.n) {
    if ((s.i & 1) == 0) s.br.refill_point();   // same cadence as the kernel: a refill point every 2 steps
    u64 rec = seq_step(t.ll, t.ml, t.of, lutLL, lutML, s);
    seqs[s.i - 1] = rec;
  }
  if (!err) err = seq_end(s, c);
  if (err) frame_fail(c, err);
}

}  // namespace zrab
