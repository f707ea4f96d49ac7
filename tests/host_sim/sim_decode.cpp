// sim_decode.cpp — host-side LOGIC TEST of the per-thread decode stages (not a product path).
//
// zra_b200/csrc/decode_core.cuh holds the thread-serial stages of the GPU decoder. The same
// source compiles as plain C++; this driver runs "one GPU thread" at a time on the CPU, in the
// same round structure as decode_kernels.cu, so tests/test_host_sim.py can check the stage logic
// against the oracle in the CPU-only test tier. The warp-cooperative executor is replaced by a
// scalar loop here (its CUDA version is covered by the -m gpu tests).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../zra_b200/csrc/decode_core.cuh"
#include "../../zra_b200/csrc/xxh64.cuh"

using namespace zrab;

// src must be readable (zero padded) up to a multiple of 4 bytes past srcBytes and 4-byte aligned.
extern "C" __attribute__((visibility("default"))) long long sim_decode_frames(const unsigned char* src, const FrameDesc* descs,
                                                                                unsigned nFrames, unsigned char* dst,
                                                                                unsigned* statusOut, unsigned* sizeOut) {
  std::vector<FrameCtx> ctx(nFrames);
  std::vector<FrameTables> tabs(nFrames);
  long long total = 0;
  for (unsigned f = 0; f < nFrames; f++) {
    const FrameDesc& d = descs[f];
    FrameCtx& c = ctx[f];
    std::vector<u8> lit(kBlockSizeMax + 16);
    u32 seqCap = kBlockSizeMax / 3 + 1;
    std::vector<u64> seqs(seqCap);
    u8* out = dst + d.dstOff;
    for (unsigned round = 0;; round++) {
      block_setup(src, d, c, tabs[f], round == 0);
      if (c.blkType == BT_NONE) break;
      if (c.blkType == BT_COMPRESSED) {
        if (c.litMode == LIT_HUF) {
          // the Huffman stage builds the two-level table from the weights (a 512-entry slot in the kernel,
          // the per-frame global area when it does not fit)
          HufLevels lv;
          if (!huf_build_two_level(tabs[f].huf, 512, tabs[f].hufWeights, c.hufCount, c.hufLog, &lv))
            huf_build_two_level(tabs[f].huf, kHufGlobalCap, tabs[f].hufWeights, c.hufCount, c.hufLog, &lv);
          for (u32 s = 0; s < c.nStreams; s++) {
            u32 e = huf_stream(src, d, c, tabs[f].huf, lv, lit.data(), s);
            if (e) frame_fail(c, e);
          }
        }
        if (c.status) break;
        seq_decode(src, d, c, tabs[f], seqs.data(), seqCap);
        if (c.status) break;
        // scalar stand-in for seq_execute
        u8* op = out + c.blkDst;
        u32 litPos = 0;
        u32 prevLit = 0, prevOut = 0;
        for (u32 i = 0; i < c.nbSeq; i++) {
          // records are cumulative: (literals consumed, bytes regenerated) after sequence i
          u32 ll = rec_lit_end(seqs[i]) - prevLit, ml = rec_out_end(seqs[i]) - prevOut - ll, off = rec_off(seqs[i]);
          prevLit = rec_lit_end(seqs[i]); prevOut = rec_out_end(seqs[i]);
          for (u32 k = 0; k < ll; k++) {
            u8 b = c.litMode == LIT_HUF ? lit[litPos + k] : (c.litMode == LIT_RAW ? src[d.srcOff + c.litSrc + litPos + k] : (u8)c.litSrc);
            op[k] = b;
          }
          op += ll; litPos += ll;
          for (u32 k = 0; k < ml; k++) op[k] = *(op + k - off);
          op += ml;
        }
        for (u32 k = litPos; k < c.litSize; k++)
          *op++ = c.litMode == LIT_HUF ? lit[k] : (c.litMode == LIT_RAW ? src[d.srcOff + c.litSrc + k] : (u8)c.litSrc);
        if ((u32)(op - (out + c.blkDst)) != c.blkOut) frame_fail(c, ZE_GENERIC);
      } else if (c.blkType == BT_RAW) {
        memcpy(out + c.blkDst, src + d.srcOff + c.blkSrc, c.blkSize);
      } else {
        memset(out + c.blkDst, src[d.srcOff + c.blkSrc], c.blkSize);
      }
    }
    if (!c.status) {
      u32 tail = (c.flags & FF_CHECKSUM) ? 4 : 0;
      if (c.srcPos + tail != d.srcLen) c.status = ZE_SRC_WRONG;
      else if (d.exact && c.dstPos != d.dstCap) c.status = ZE_CORRUPTION;
      else if (c.fcs != ~0ull && c.fcs != c.dstPos) c.status = ZE_CORRUPTION;
      else if (c.flags & FF_CHECKSUM) {
        // same quad-lane structure as k_frame_finish
        u32 len = c.dstPos;
        u64 acc[4];
        for (u32 q = 0; q < 4; q++) {
          acc[q] = xxh_init_acc(q);
          for (u32 k = 0; k < (len >> 5); k++) acc[q] = xxh_round(acc[q], ld64(out + 32 * (u64)k + 8 * q));
        }
        u64 h;
        if (len >= 32) {
          h = xxh_rotl(acc[0], 1) + xxh_rotl(acc[1], 7) + xxh_rotl(acc[2], 12) + xxh_rotl(acc[3], 18);
          for (u32 q = 0; q < 4; q++) h = xxh_merge(h, acc[q]);
        } else {
          h = kXP5;
        }
        h = xxh_finish(h, len, out + (len & ~31u), len & 31u);
        if ((u32)h != ld32(src + d.srcOff + c.srcPos)) c.status = ZE_CHECKSUM_WRONG;
      }
    }
    statusOut[f] = c.status;
    sizeOut[f] = c.dstPos;
    total += c.dstPos;
  }
  return total;
}
