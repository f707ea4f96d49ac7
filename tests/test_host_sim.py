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
