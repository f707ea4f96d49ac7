"""The reference's own CLI (programs/zratool.cpp, UNMODIFIED) as an acceptance test.

oracle/Makefile compiles that one source twice, in place from /root/reference: `_ref/ZraTool_ref` against the reference
library and `_ref/ZraTool_b200` against this repo's include/zra.hpp + zra_b200/libzra_b200.so. That it compiles and links
at all is the drop-in check of the C++ API surface (zratool.cpp:127-279 uses Compressor, FullDecompressor, Decompressor,
CompressBuffer, DecompressBuffer, DecompressRA and Header fields). The CPU tests cover BASELINE configs[0]'s shape (text,
16 KiB frames, level 3, streaming and in-memory modes) on the reference binary; the GPU tests cross the two binaries."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from zra_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "ZraTool_ref")
GPU_TOOL = os.path.join(ROOT, "oracle", "_ref", "ZraTool_b200")

need_ref = pytest.mark.skipif(not os.path.exists(REF_TOOL), reason="oracle/_ref/ZraTool_ref not built (needs /root/reference)")
need_gpu_tool = pytest.mark.skipif(not os.path.exists(GPU_TOOL), reason="oracle/_ref/ZraTool_b200 not built (needs /root/reference)")


def run(tool, *args, ok=True):
    r = subprocess.run([tool, *map(str, args)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    if ok:
        assert r.returncode == 0, (r.returncode, r.stderr[-800:])
    return r


def make_input(tmp_path, size, seed=7, name="in.bin"):
    data = synth.text(size, seed=seed)
    p = tmp_path / name
    data.tofile(p)
    return p, data


@need_ref
@pytest.mark.parametrize("mode_c,mode_d", [("c", "d"), ("imc", "imd"), ("c", "imd"), ("imc", "d")])
def test_reference_cli_round_trip(tmp_path, mode_c, mode_d):
    """configs[0] shape on the reference path (CPU): text, 16 KiB frames, level 3; also checked against the C oracle."""
    import refzra

    src, data = make_input(tmp_path, (4 << 20) + 12345)
    run(REF_TOOL, mode_c, src, 3, 16384)
    arc = np.fromfile(str(src) + ".zra", dtype=np.uint8)
    assert np.array_equal(refzra.oracle_decompress_buffer(arc), data)
    os.remove(src)
    run(REF_TOOL, mode_d, str(src) + ".zra")
    assert np.array_equal(np.fromfile(src, dtype=np.uint8), data)


@need_gpu_tool
def test_gpu_cli_fails_loudly_without_a_device(tmp_path):
    """No CPU fallback: without a CUDA device the drop-in CLI must not produce an archive."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    src, _ = make_input(tmp_path, 200000)
    r = run(GPU_TOOL, "imc", src, 3, 16384, ok=False)
    assert r.returncode != 0
    assert "no usable CUDA device" in r.stderr
    assert not os.path.exists(str(src) + ".zra") or os.path.getsize(str(src) + ".zra") == 0


@pytest.mark.gpu
@need_ref
@need_gpu_tool
@pytest.mark.parametrize("frame,level", [(16384, 3), (65536, 1), (262144, 3)])
@pytest.mark.parametrize("mode_c,mode_d", [("c", "d"), ("imc", "imd")])
def test_cli_cross_round_trips(tmp_path, frame, level, mode_c, mode_d):
    """GPU-written archives decode bit-exactly with the reference CLI and vice versa, streaming and in-memory."""
    src, data = make_input(tmp_path, (6 << 20) + 4321)
    # GPU compress -> reference decompress
    run(GPU_TOOL, mode_c, src, level, frame, 2)
    shutil.move(str(src) + ".zra", tmp_path / "gpu.zra")
    os.remove(src)
    run(REF_TOOL, mode_d, tmp_path / "gpu.zra", 2)
    assert np.array_equal(np.fromfile(tmp_path / "gpu", dtype=np.uint8), data)
    # reference compress -> GPU decompress
    data.tofile(src)
    run(REF_TOOL, mode_c, src, level, frame, 2)
    shutil.move(str(src) + ".zra", tmp_path / "ref.zra")
    run(GPU_TOOL, mode_d, tmp_path / "ref.zra", 2)
    assert np.array_equal(np.fromfile(tmp_path / "ref", dtype=np.uint8), data)
    # both writers agree on the header geometry (sizes differ: the frames are not byte-identical)
    from common import parse_header

    hg = parse_header(np.fromfile(tmp_path / "gpu.zra", dtype=np.uint8))
    hr = parse_header(np.fromfile(tmp_path / "ref.zra", dtype=np.uint8))
    for k in ("frameId", "headerSize", "magic", "version", "uncompressedSize", "tableSize", "frameSize", "metaSize"):
        assert hg[k] == hr[k], k


@pytest.mark.gpu
@need_gpu_tool
def test_cli_benchmark_mode(tmp_path):
    """`b` mode: in-memory + streaming compress/decompress and both random-access paths with zratool's own memcmp check."""
    src, _ = make_input(tmp_path, 3 << 20)
    r = run(GPU_TOOL, "b", src, 3, 16384, 1, 4096, 65536)
    assert "In-Memory RA Summary" in r.stdout and "Streaming RA Summary" in r.stdout
