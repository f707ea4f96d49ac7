"""CPU-tier logic test of the GPU decoder's thread-serial stages (zra_b200/csrc/decode_core.cuh
compiled as plain C++ by tests/host_sim) against the golden vectors and the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import refzra
from common import golden_archive, golden_archives, golden_frame, golden_frames, parse_header, seek_table, sha

HERE = os.path.dirname(os.path.abspath(__file__))


class FD(C.Structure):
    _fields_ = [("srcOff", C.c_uint64), ("dstOff", C.c_uint64), ("srcLen", C.c_uint32), ("dstCap", C.c_uint32),
                ("exact", C.c_uint32), ("pad", C.c_uint32)]


@pytest.fixture(scope="module")
def sim():
    L = C.CDLL(os.path.join(HERE, "host_sim", "libsim_decode.so"))
    L.sim_decode_frames.restype = C.c_longlong
    return L


def run(sim, src, frames):
    pad = np.zeros(((src.size + 3) // 4) * 4 + 16, np.uint8)
    pad[: src.size] = src
    n = len(frames)
    arr = (FD * n)()
    total = 0
    for i, (so, sl, do, dc, ex) in enumerate(frames):
        arr[i] = FD(so, do, sl, dc, ex, 0)
        total = max(total, do + dc)
    dst = np.full(total + 64, 0xAA, np.uint8)
    st = (C.c_uint32 * max(n, 1))()
    sz = (C.c_uint32 * max(n, 1))()
    sim.sim_decode_frames(pad.ctypes.data_as(C.c_void_p), arr, n, dst.ctypes.data_as(C.c_void_p), st, sz)
    return dst, list(st)[:n], list(sz)[:n]


@pytest.fixture(scope="module")
def simk():
    """The real k_seq_decode / k_seq_execute kernels under the SIMT emulator (tests/host_sim/simt.h)."""
    L = C.CDLL(os.path.join(HERE, "host_sim", "libsim_kernels.so"))
    L.sim_kernels_decode.restype = C.c_longlong
    return L


def run_kernels(simk, src, frames, mode=0, guard=4096):
    pad = np.zeros(((src.size + 15) // 16) * 16 + 64, np.uint8)
    pad[: src.size] = src
    n = len(frames)
    arr = (FD * n)()
    total, biggest = 0, 1
    for i, (so, sl, do, dc, ex) in enumerate(frames):
        arr[i] = FD(so, do, sl, dc, ex, 0)
        total = max(total, do + dc)
        biggest = max(biggest, dc)
    dst = np.full(total + guard, 0xAA, np.uint8)
    st = (C.c_uint32 * max(n, 1))()
    sz = (C.c_uint32 * max(n, 1))()
    simk.sim_kernels_decode(pad.ctypes.data_as(C.c_void_p), arr, n, dst.ctypes.data_as(C.c_void_p), biggest, st, sz, mode)
    assert (dst[total:] == 0xAA).all(), "wrote past the end of the destination"
    return dst, list(st)[:n], list(sz)[:n]


def crafted_oversized_block(nseq):
    """a 31-byte frame whose single block regenerates nseq * 65540 bytes (RLE tables, ML code 52,
    repeat offset 1): above Block_Maximum_Size, and above the 18-bit record fields from nseq = 4."""
    lits = bytes([(nseq << 3) | 0]) + b"a" * nseq                                        # raw literals, one per sequence
    seqs = bytes([nseq, 0x54, 1, 0, 52]) + b"\x00" * (2 * nseq) + b"\x01"                 # RLE LL=1, OF=0, ML=52; extras all 0
    content = lits + seqs
    bh = (len(content) << 3) | (2 << 1) | 1
    return np.frombuffer(b"\x28\xb5\x2f\xfd" + b"\x00\x50" + bytes([bh & 255, (bh >> 8) & 255, bh >> 16]) + content, np.uint8).copy()


def crafted_huge_literals():
    """a tiny frame whose block declares 128 KiB - 1 Huffman literals (5-byte literals header)."""
    lit_size, comp = (1 << 17) - 1, 40
    w = 2 | (3 << 2) | (lit_size << 4) | (comp << 22)
    hdr = bytes([(w >> (8 * i)) & 255 for i in range(4)]) + bytes([comp >> 10])
    content = hdr + bytes(range(1, comp + 1)) + b"\x00"
    bh = (len(content) << 3) | (2 << 1) | 1
    return np.frombuffer(b"\x28\xb5\x2f\xfd" + b"\x00\x50" + bytes([bh & 255, (bh >> 8) & 255, bh >> 16]) + content, np.uint8).copy()


def archive_frames(archive):
    h = parse_header(archive)
    t = seek_table(archive)
    fs, U = h["frameSize"], h["uncompressedSize"]
    return [(h["size"] + int(t[i]), int(t[i + 1] - t[i]), i * fs, min(fs, U - i * fs), 1) for i in range(len(t) - 1)]


@pytest.mark.parametrize("name", golden_archives())
def test_archives(sim, name):
    archive, meta = golden_archive(name)
    dst, st, sz = run(sim, archive, archive_frames(archive))
    assert not any(st), st
    assert sha(dst[: meta["bytes"]]) == meta["sha256"]


def test_decodecorpus(sim):
    bad = []
    for name in golden_frames():
        z, meta = golden_frame(name)
        dst, st, sz = run(sim, z, [(0, z.size, 0, meta["bytes"], 0)])
        if st[0] or sz[0] != meta["bytes"] or sha(dst[: meta["bytes"]]) != meta["sha256"]:
            bad.append((name, st[0], sz[0]))
    assert not bad, bad


def test_corruption_codes_agree_with_oracle(sim):
    archive, _ = golden_archive("text_f16384_l3")
    frames = archive_frames(archive)
    h = parse_header(archive)
    rng = np.random.default_rng(3)
    for _ in range(80):
        bad = archive.copy()
        pos = int(rng.integers(h["size"], archive.size))
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        _, st, _ = run(sim, bad, frames)
        sim_err = next((s for s in st if s), 0)
        try:
            refzra.oracle_decompress_buffer(bad)
            ora = 0
        except refzra.OracleError as e:
            ora = e.zstd
        assert bool(sim_err) == bool(ora), (pos, sim_err, ora)


# ---- the warp-cooperative kernels themselves, under the SIMT emulator
@pytest.mark.parametrize("name", golden_archives())
def test_kernels_archives(simk, name):
    archive, meta = golden_archive(name)
    frames = archive_frames(archive)[:12]   # the emulator runs ~1 frame / 0.2 s
    expect = refzra.oracle_decompress_buffer(archive)
    dst, st, sz = run_kernels(simk, archive, frames)
    assert not any(st), st
    n = sum(f[3] for f in frames)
    assert np.array_equal(dst[:n], expect[:n])


@pytest.mark.parametrize("mode", [0, 1])
def test_kernels_decodecorpus(simk, mode):
    bad = []
    for name in golden_frames()[:: 2 if mode else 1]:
        z, meta = golden_frame(name)
        dst, st, sz = run_kernels(simk, z, [(0, z.size, 0, meta["bytes"], 0)], mode)
        if st[0] or sz[0] != meta["bytes"] or sha(dst[: meta["bytes"]]) != meta["sha256"]:
            bad.append((name, st[0], sz[0]))
    assert not bad, bad


def test_kernels_many_frames_per_lane(simk):
    """More frames than table slots: lanes pull several frames, finish at different steps."""
    archive, meta = golden_archive("text_f1000_l5")
    frames = archive_frames(archive)
    dst, st, sz = run_kernels(simk, archive, frames)
    assert not any(st)
    assert sha(dst[: meta["bytes"]]) == meta["sha256"]


def test_kernels_corruption_codes_agree_with_serial_path(sim, simk):
    """The fast sequence loop hands every frame it flags to the careful path: same status as the thread-serial decoder."""
    archive, _ = golden_archive("text_f16384_l3")
    frames = archive_frames(archive)[:6]
    h = parse_header(archive)
    end = frames[-1][0] + frames[-1][1]
    rng = np.random.default_rng(5)
    for _ in range(60):
        bad = archive.copy()
        pos = int(rng.integers(h["size"], end))
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        _, st_serial, _ = run(sim, bad, frames)
        _, st_kern, _ = run_kernels(simk, bad, frames)
        assert st_serial == st_kern, (pos, st_serial, st_kern)


@pytest.mark.parametrize("nseq", [2, 3, 4, 5, 7])
def test_crafted_block_above_block_maximum_is_refused(sim, simk, nseq):
    z = crafted_oversized_block(nseq)
    cap = 1 << 20
    dst, st, sz = run_kernels(simk, z, [(0, z.size, 0, cap, 0)])
    assert st[0] == 70, st          # dstSize_tooSmall: the block may not regenerate more than 128 KiB
    _, st2, _ = run(sim, z, [(0, z.size, 0, cap, 0)])
    assert st2[0] == 70
    # the same frame with ONE sequence (65 540 bytes) is legal and decodes
    z1 = crafted_oversized_block(1)
    dst, st, sz = run_kernels(simk, z1, [(0, z1.size, 0, cap, 0)])
    assert st[0] == 0 and sz[0] == 65540 and (dst[:65540] == ord("a")).all()


def test_crafted_literals_larger_than_the_frame_are_refused(sim, simk):
    z = crafted_huge_literals()
    for cap in (16, 4096, 65536):
        _, st, _ = run_kernels(simk, z, [(0, z.size, 0, cap, 0)])
        assert st[0] == 70, (cap, st)
        _, st2, _ = run(sim, z, [(0, z.size, 0, cap, 0)])
        assert st2[0] == 70
