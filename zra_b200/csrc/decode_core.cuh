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
  u32 pad;
};

// Table scratch of one frame (HBM). The sequence tables use the compact 2-byte format.
struct FrameTables {
  CSym ll[512];
  CSym ml[512];
  CSym of[256];
  HufSym huf[4096];
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
ZRA_DEV void block_setup(const u8* srcBase, const FrameDesc& d, FrameCtx& c, FrameTables& t, bool firstRound) {
  const u8* f = srcBase + d.srcOff;
  if (firstRound) { if (!parse_frame_header(f, d, c)) return; }
  c.blkType = BT_NONE;
  if (c.status || (c.flags & FF_DONE)) return;
  u32 tail = (c.flags & FF_CHECKSUM) ? 4u : 0u;
  (void)tail;
  if (c.srcPos + 3 > d.srcLen) { frame_fail(c, ZE_SRC_WRONG); return; }
  u32 bh = ld24(f + c.srcPos);
  u32 last = bh & 1, type = (bh >> 1) & 3, bsz = bh >> 3;
  if (type == 3) { frame_fail(c, ZE_CORRUPTION); return; }
  u32 csz = (type == BT_RLE) ? 1 : bsz;
  u32 content = c.srcPos + 3;
  if (content + csz > d.srcLen) { frame_fail(c, ZE_SRC_WRONG); return; }
  c.blkSrc = content;
  c.blkSize = bsz;
  c.blkDst = c.dstPos;
  c.srcPos = content + csz;
  if (last) c.flags |= FF_DONE;
  if (type != BT_COMPRESSED) {
    if (bsz > d.dstCap - c.dstPos) { frame_fail(c, ZE_DST_TOO_SMALL); return; }
    c.blkType = type;
    c.blkOut = bsz;
    c.dstPos += bsz;
    return;
  }
  // ---- compressed block
  if (csz >= kBlockSizeMax) { frame_fail(c, ZE_SRC_WRONG); return; }
  if (csz < 3) { frame_fail(c, ZE_CORRUPTION); return; }
  const u8* b = f + content;
  u32 used;
  c.nStreams = 0;
  {
    u32 ltype = b[0] & 3, fmt = (b[0] >> 2) & 3;
    if (ltype >= 2) {
      u32 lh, cs, litSize;
      bool single = false;
      if (ltype == 3 && !(c.flags & FF_HUF_VALID)) { frame_fail(c, ZE_DICT_CORRUPTED); return; }
      if (csz < 5) { frame_fail(c, ZE_CORRUPTION); return; }
      u32 w = ld32(b);
      if (fmt <= 1) { single = !fmt; lh = 3; litSize = (w >> 4) & 0x3FF; cs = (w >> 14) & 0x3FF; }
      else if (fmt == 2) { lh = 4; litSize = (w >> 4) & 0x3FFF; cs = w >> 18; }
      else { lh = 5; litSize = (w >> 4) & 0x3FFFF; cs = (w >> 22) + ((u32)b[4] << 10); }
      if (litSize > kBlockSizeMax || cs + lh > csz) { frame_fail(c, ZE_CORRUPTION); return; }
      u32 hoff = content + lh, hlen = cs;
      if (ltype == 2) {
        u8 weights[256];
        u32 count, log;
        u32 h = huf_read_weights(srcBase, d.srcOff + hoff, hlen, weights, &count, &log);
        if (!h || !huf_build_table(t.huf, weights, count, log)) { frame_fail(c, ZE_CORRUPTION); return; }
        c.hufLog = log;
        c.flags |= FF_HUF_VALID;
        hoff += h; hlen -= h;
      }
      if (single) {
        c.nStreams = 1;
        c.strOff[0] = hoff; c.strLen[0] = hlen;
      } else {
        if (hlen < 10) { frame_fail(c, ZE_CORRUPTION); return; }
        const u8* j = f + hoff;
        u32 l1 = ld16(j), l2 = ld16(j + 2), l3 = ld16(j + 4);
        if (l1 + l2 + l3 + 6 > hlen) { frame_fail(c, ZE_CORRUPTION); return; }
        u32 seg = (litSize + 3) / 4;
        if (seg * 3 > litSize) { frame_fail(c, ZE_CORRUPTION); return; }
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
        if (lh + litSize > csz) { frame_fail(c, ZE_CORRUPTION); return; }
        c.litMode = LIT_RAW; c.litSrc = content + lh; c.litSize = litSize;
        used = lh + litSize;
      } else {
        if (lh + 1 > csz || litSize > kBlockSizeMax) { frame_fail(c, ZE_CORRUPTION); return; }
        c.litMode = LIT_RLE; c.litSrc = b[lh]; c.litSize = litSize;
        used = lh + 1;
      }
    }
  }
  // ---- sequences section header
  const u8* ip = b + used;
  const u8* iend = b + csz;
  if (ip >= iend) { frame_fail(c, ZE_SRC_WRONG); return; }
  u32 nbSeq = *ip++;
  if (!nbSeq) {
    if (ip != iend) { frame_fail(c, ZE_SRC_WRONG); return; }
  } else {
    if (nbSeq > 0x7F) {
      if (nbSeq == 0xFF) {
        if (ip + 2 > iend) { frame_fail(c, ZE_SRC_WRONG); return; }
        nbSeq = ld16(ip) + kLongNbSeq; ip += 2;
      } else {
        if (ip >= iend) { frame_fail(c, ZE_SRC_WRONG); return; }
        nbSeq = ((nbSeq - 0x80) << 8) + *ip++;
      }
    }
    if (ip + 1 > iend) { frame_fail(c, ZE_SRC_WRONG); return; }
    u32 modes = *ip++;
    bool rep = (c.flags & FF_FSE_VALID) != 0;
    u32 h = setup_seq_table(t.ll, &c.llLog, modes >> 6, SEQ_LL, ip, (u32)(iend - ip), rep);
    if (h == 0xFFFFFFFFu) { frame_fail(c, ZE_CORRUPTION); return; }
    ip += h;
    h = setup_seq_table(t.of, &c.ofLog, (modes >> 4) & 3, SEQ_OF, ip, (u32)(iend - ip), rep);
    if (h == 0xFFFFFFFFu) { frame_fail(c, ZE_CORRUPTION); return; }
    ip += h;
    h = setup_seq_table(t.ml, &c.mlLog, (modes >> 2) & 3, SEQ_ML, ip, (u32)(iend - ip), rep);
    if (h == 0xFFFFFFFFu) { frame_fail(c, ZE_CORRUPTION); return; }
    ip += h;
    c.flags |= FF_FSE_VALID;
    if (ip >= iend) { frame_fail(c, ZE_CORRUPTION); return; }  // the bitstream needs at least its end mark
  }
  c.nbSeq = nbSeq;
  c.seqOff = (u32)(ip - f);
  c.seqLen = (u32)(iend - ip);
  c.blkType = BT_COMPRESSED;
  c.blkOut = 0;
  if (!nbSeq) {  // literals only: nothing for the sequence stage to do
    if (c.litSize > d.dstCap - c.blkDst) { frame_fail(c, ZE_DST_TOO_SMALL); return; }
    c.blkOut = c.litSize;
    c.dstPos = c.blkDst + c.litSize;
  }
}

// ---------------------------------------------------------------- huf_stream (1 thread / stream)
// Decodes stream `s` of the current block of one frame into the frame's literal scratch.
// Returns 0 or a ZErr.
ZRA_DEV u32 huf_stream(const u8* srcBase, const FrameDesc& d, const FrameCtx& c, const HufSym* table, u8* lit, u32 s) {
  u32 seg = c.nStreams == 4 ? (c.litSize + 3) / 4 : c.litSize;
  u32 outBeg = s * seg;
  u32 n = (c.nStreams == 4 && s == 3) ? c.litSize - 3 * seg : seg;
  BackReader br;
  if (!br.init(srcBase, d.srcOff + c.strOff[s], c.strLen[s])) return ZE_CORRUPTION;
  const u32 log = c.hufLog;
  u8* out = lit + outBeg;
  for (u32 i = 0; i < n; i++) {
    br.refill();
    HufSym e = table[br.peek(log)];
    br.skip(e >> 8);
    out[i] = (u8)e;
  }
  return br.remaining == 0 ? ZE_OK : ZE_CORRUPTION;
}

// ---------------------------------------------------------------- seq_decode (1 thread / frame)
// Packed sequence record: ll (18 bits) | ml (18 bits) << 18 | offset (28 bits) << 36.
ZRA_DEV u64 seq_pack(u32 ll, u32 ml, u32 off) { return (u64)ll | ((u64)ml << 18) | ((u64)off << 36); }
ZRA_DEV u32 seq_ll(u64 s) { return (u32)s & 0x3FFFFu; }
ZRA_DEV u32 seq_ml(u64 s) { return (u32)(s >> 18) & 0x3FFFFu; }
ZRA_DEV u32 seq_off(u64 s) { return (u32)(s >> 36); }
constexpr u32 kMaxOffset = (1u << 28) - 1;

// Per-frame state of the sequence stage; lives in registers of the lane that owns the frame.
struct SeqState {
  BackReader br;
  u32 sLL, sML, sOF;       // FSE states
  u32 rep0, rep1, rep2;    // repeat-offset history
  u32 litUsed, produced;   // literals consumed / bytes regenerated so far in this block
  u32 i, n;                // sequence cursor / count
  u32 llLog, ofLog, mlLog;
  u32 litSize, room, blkDst;
};

// Starts the sequence stage of the current block of one frame. Returns 0 or a ZErr.
ZRA_DEV u32 seq_begin(const u8* srcBase, const FrameDesc& d, const FrameCtx& c, u32 seqCap, SeqState& s) {
  if (c.nbSeq > seqCap) return ZE_CORRUPTION;
  if (!s.br.init(srcBase, d.srcOff + c.seqOff, c.seqLen)) return ZE_CORRUPTION;
  s.llLog = c.llLog; s.ofLog = c.ofLog; s.mlLog = c.mlLog;
  s.sLL = s.br.read(s.llLog);
  s.sOF = s.br.read(s.ofLog);
  s.br.refill();
  s.sML = s.br.read(s.mlLog);
  s.rep0 = c.rep[0]; s.rep1 = c.rep[1]; s.rep2 = c.rep[2];
  s.litUsed = 0; s.produced = 0; s.i = 0; s.n = c.nbSeq;
  s.litSize = c.litSize; s.room = d.dstCap - c.blkDst; s.blkDst = c.blkDst;
  return ZE_OK;
}

// Decodes and validates ONE sequence. tLL/tML/tOF are the compact tables (shared memory in the
// kernel), lutLL/lutML the packed baseline tables. Returns 0 or a ZErr; *rec receives the record.
ZRA_DEV u32 seq_step(const CSym* tLL, const CSym* tML, const CSym* tOF, const u32* lutLL, const u32* lutML, SeqState& s, u64* rec) {
  BackReader& br = s.br;
  const u32 eLL = tLL[s.sLL], eML = tML[s.sML], eOF = tOF[s.sOF];
  const u32 lutl = lutLL[csym_symbol((CSym)eLL)], lutm = lutML[csym_symbol((CSym)eML)];
  const u32 ofBits = csym_symbol((CSym)eOF);
  const u32 llBase = lutl & 0xFFFFFFu, llBits = lutl >> 24;
  const u32 mlBase = lutm & 0xFFFFFFu, mlBits = lutm >> 24;
  u32 offset;
  br.refill();
  if (ofBits > 1) {
    offset = of_base(ofBits) + br.read(ofBits);
    s.rep2 = s.rep1; s.rep1 = s.rep0; s.rep0 = offset;
  } else {
    u32 ll0 = (llBase == 0);
    if (ofBits == 0) {
      if (!ll0) offset = s.rep0;
      else { offset = s.rep1; s.rep1 = s.rep0; s.rep0 = offset; }
    } else {
      u32 idx = 1 + ll0 + br.read(1);
      u32 v = (idx == 3) ? s.rep0 - 1 : (idx == 1 ? s.rep1 : s.rep2);
      v += !v;
      if (idx != 1) s.rep2 = s.rep1;
      s.rep1 = s.rep0; s.rep0 = offset = v;
    }
  }
  br.refill();
  const u32 ml = mlBase + br.read(mlBits);
  const u32 ll = llBase + br.read(llBits);
  br.refill();
  if (s.i + 1 < s.n) {  // the last sequence leaves the states alone
    u32 ns = csym_ns((CSym)eLL), nb = s.llLog - highbit32(ns);
    s.sLL = (ns << nb) - (1u << s.llLog) + br.read(nb);
    ns = csym_ns((CSym)eML); nb = s.mlLog - highbit32(ns);
    s.sML = (ns << nb) - (1u << s.mlLog) + br.read(nb);
    ns = csym_ns((CSym)eOF); nb = s.ofLog - highbit32(ns);
    s.sOF = (ns << nb) - (1u << s.ofLog) + br.read(nb);
  }
  s.i++;
  // validation: everything seq_execute will trust
  if (ll > s.litSize - s.litUsed) return ZE_CORRUPTION;
  s.litUsed += ll;
  if (ll + ml > s.room - s.produced) return ZE_DST_TOO_SMALL;  // produced <= room is an invariant
  if (offset > s.blkDst + s.produced + ll || offset > kMaxOffset) return ZE_CORRUPTION;
  s.produced += ll + ml;
  *rec = seq_pack(ll, ml, offset);
  return ZE_OK;
}

// Finishes the block after its last sequence: stream exhaustion, trailing literals, write-back.
ZRA_DEV u32 seq_end(const SeqState& s, FrameCtx& c) {
  if (s.br.remaining != 0) return ZE_CORRUPTION;
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
  u32 lutLL[36], lutML[53];
  for (u32 k = 0; k < 36; k++) lutLL[k] = ll_lut(k);
  for (u32 k = 0; k < 53; k++) lutML[k] = ml_lut(k);
  SeqState s;
  u32 err = seq_begin(srcBase, d, c, seqCap, s);
  while (!err && s.i < s.n) {
    u64 rec;
    err = seq_step(t.ll, t.ml, t.of, lutLL, lutML, s, &rec);
    if (!err) seqs[s.i - 1] = rec;
  }
  if (!err) err = seq_end(s, c);
  if (err) frame_fail(c, err);
}

}  // namespace zrab
