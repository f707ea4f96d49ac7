// sim_encode.cpp — host-side LOGIC TEST of the per-thread ENCODER stages (not a product path).
//
// zra_b200/csrc/enc_core.cuh holds the thread-serial stages of the GPU encoder; compiled as plain
// C++ here, one "GPU thread" at a time in the same round structure as encode_kernels.cu, so that
// tests/test_host_sim.py can check (CPU tier) that the frames it writes are valid zstd — decodable
// by the oracle, the reference and stock libzstd — and how their size compares with the reference's.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../zra_b200/csrc/enc_core.cuh"

using namespace zrab;

// Encodes src[0..n) (4-byte aligned, readable to n+8) as ONE zstd frame into dst; returns its size.
extern "C" __attribute__((visibility("default"))) long long sim_encode_frame(const unsigned char* src, unsigned n, int level, int checksum,
                                                                               unsigned char* dst, unsigned dstCap) {
  EncParams p = enc_params(level, n, checksum != 0);
  std::vector<u32> tabS(1u << p.hashLogS, 0), tabL(p.hashLogL ? (1u << p.hashLogL) : 1, 0);
  const u32 blk = n < kBlockSizeMax ? n : kBlockSizeMax;
  std::vector<u64> seqs(blk / 3 + 2);
  std::vector<u8> lit(blk + 16), hufOut(4 * (size_t)(blk + 64)), hdr(512), seqOut(blk * 2 + 64), cells(1024);
  std::vector<u32> hist(256);
  std::vector<HufCode> hcodes(256);
  std::vector<FseSymTT> tt(36 + 32 + 53);
  std::vector<u16> states(512 + 256 + 512);
  EncScratch s;
  s.tabS = tabS.data(); s.tabL = tabL.data(); s.seqs = seqs.data(); s.lit = lit.data(); s.hist = hist.data();
  s.hcodes = hcodes.data(); s.hufOut = hufOut.data(); s.hufStride = blk + 64; s.hdr = hdr.data(); s.tt = tt.data();
  s.states = states.data(); s.seqOut = seqOut.data(); s.seqOutCap = (u32)seqOut.size(); s.cells = cells.data(); s.cnt = nullptr;
  EncCtx c;
  memset(&c, 0, sizeof(c));
  c.srcLen = n;
  c.rep[0] = 1; c.rep[1] = 4; c.rep[2] = 8;
  c.windowLog = p.windowLogMax;
  std::vector<u8> out((size_t)n + (n >> 7) + 1024);
  for (u32 pos = 0; pos < n || pos == 0; pos += kBlockSizeMax) {
    c.blkPos = pos;
    c.blkLen = n - pos < kBlockSizeMax ? n - pos : kBlockSizeMax;
    c.lastBlock = pos + c.blkLen >= n;
    enc_match(src, 0, p, c, s);
    enc_literals(src, 0, c, s);
    enc_plan(c, s);
    for (u32 st = 0; st < 4; st++) enc_huf(c, s, st);
    enc_seq(c, s);
    enc_assemble(src, 0, p, c, s, out.data());
    if (c.lastBlock) break;
  }
  if (p.checksum) {
    u32 h = enc_checksum_serial(src, 0, n);
    for (int k = 0; k < 4; k++) out[c.outPos++] = (u8)(h >> (8 * k));
  }
  if (c.outPos > dstCap) return -1;
  memcpy(dst, out.data(), c.outPos);
  return c.outPos;
}
