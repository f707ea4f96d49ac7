// sim_kernels.cpp — host-side LOGIC TEST of the warp-cooperative decode kernels (not a product path).
//
// k_seq_decode (decode_seq.cuh) and k_seq_execute (decode_exec.cuh) are compiled unchanged by g++ and run under the
// SIMT emulator (simt.h: one fiber per CUDA thread, collectives rendezvous like on the GPU). The other stages run as
// the thread-serial functions of decode_core.cuh, in the round structure of decode_kernels.cu.
#include "simt.h"

#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../zra_b200/csrc/decode_core.cuh"
#include "../../zra_b200/csrc/decode_exec.cuh"
#include "../../zra_b200/csrc/decode_seq.cuh"
#include "../../zra_b200/csrc/xxh64.cuh"

using namespace zrab;

// src must be readable (zero padded) 64 bytes past its end and 16-byte aligned.
// mode bit 0: force the general sequence geometry (no small-table split).
extern "C" __attribute__((visibility("default"))) long long sim_kernels_decode(const unsigned char* src, const FrameDesc* descs,
                                                                                 unsigned nFrames, unsigned char* dst, unsigned maxDstCap,
                                                                                 unsigned* statusOut, unsigned* sizeOut, unsigned mode) {
  const u32 blk = maxDstCap < kBlockSizeMax ? maxDstCap : kBlockSizeMax;
  const u32 litStride = ((blk + 15u) & ~15u) ? ((blk + 15u) & ~15u) : 16u;
  const u32 seqStride = blk / 3 + 1;
  std::vector<FrameCtx> ctx(nFrames);
  std::vector<FrameTables> tabs(nFrames);
  std::vector<u8> lit((size_t)litStride * nFrames + 64);
  std::vector<u64> seqs((size_t)seqStride * nFrames);
  std::vector<u32> seqList(nFrames), redoList(nFrames);
  const bool splitSmall = litStride <= (32u << 10) && !(mode & 1);
  for (unsigned round = 0; round < 4096; round++) {
    RoundWork work;
    memset(&work, 0, sizeof(work));
    unsigned liveFrames = 0;
    for (unsigned f = 0; f < nFrames; f++) {
      FrameCtx& c = ctx[f];
      block_setup(src, descs[f], c, tabs[f], round == 0);
      if (c.blkType == BT_NONE) continue;
      liveFrames++;
      if (c.blkType != BT_COMPRESSED || c.status) continue;
      if (c.litMode == LIT_HUF && c.litSize) {
        HufLevels lv;
        if (!huf_build_two_level(tabs[f].huf, 512, tabs[f].hufWeights, c.hufCount, c.hufLog, &lv))
          huf_build_two_level(tabs[f].huf, kHufGlobalCap, tabs[f].hufWeights, c.hufCount, c.hufLog, &lv);
        for (u32 s = 0; s < c.nStreams; s++) {
          u32 e = huf_stream(src, descs[f], c, tabs[f].huf, lv, lit.data() + (size_t)f * litStride, s);
          if (e && !c.status) c.status = e;
        }
      }
      if (c.nbSeq) {
        if (splitSmall && c.llLog <= kSeqSmallLogMax && c.mlLog <= kSeqSmallLogMax) seqList[nFrames - 1 - work.seqCountS++] = f;
        else seqList[work.seqCount++] = f;
      }
    }
    if (!liveFrames) break;
    const u32 ctas = (nFrames + 63) / 64 ? (nFrames + 63) / 64 : 1;  // fewer slots than frames: lanes pull several frames
    if (splitSmall)
      simt::launch(simt::Dim3(ctas), simt::Dim3(32), SeqGeom<true>::kSmem, [&] {
        k_seq_decode<true>(src, descs, ctx.data(), tabs.data(), seqs.data(), seqStride, &work, seqList.data(), redoList.data(), nFrames);
      });
    simt::launch(simt::Dim3(ctas), simt::Dim3(32), SeqGeom<false>::kSmem, [&] {
      k_seq_decode<false>(src, descs, ctx.data(), tabs.data(), seqs.data(), seqStride, &work, seqList.data(), redoList.data(), nFrames);
    });
    for (u32 i = 0; i < work.redoCount; i++) {
      const u32 f = redoList[i];
      seq_decode(src, descs[f], ctx[f], tabs[f], seqs.data() + (size_t)f * seqStride, seqStride);
    }
    simt::launch(simt::Dim3((nFrames + kExecWarps - 1) / kExecWarps), simt::Dim3(kExecWarps * 32), 0, [&] {
      k_seq_execute(src, dst, descs, ctx.data(), lit.data(), litStride, seqs.data(), seqStride, nFrames);
    });
  }
  long long total = 0;
  for (unsigned f = 0; f < nFrames; f++) {
    FrameCtx& c = ctx[f];
    const FrameDesc& d = descs[f];
    u8* out = dst + d.dstOff;
    if (!c.status) {
      u32 tail = (c.flags & FF_CHECKSUM) ? 4 : 0;
      if (c.srcPos + tail != d.srcLen) c.status = ZE_SRC_WRONG;
      else if (d.exact && c.dstPos != d.dstCap) c.status = ZE_CORRUPTION;
      else if (c.fcs != ~0ull && c.fcs != c.dstPos) c.status = ZE_CORRUPTION;
      else if (c.flags & FF_CHECKSUM) {
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
