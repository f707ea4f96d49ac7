"""CPU tier: the C-ABI library loads, exports every symbol the headers declare, its GPU-free host
logic matches the oracle, and GPU entry points FAIL LOUDLY (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import refzra
import zra_b200
from common import golden_archive, parse_header

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in ("zra.h", "zra_b200.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        names += re.findall(r"ZRA_EXPORT[^;(]*?\b(Zra\w+)\s*\(", text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    L = C.CDLL(zra_b200.lib_path())
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_cpp_api_symbols_present():
    import subprocess

    out = subprocess.run(["nm", "-DC", "--defined-only", zra_b200.lib_path()], capture_output=True, text=True).stdout
    for sym in ("zra::CompressBuffer(", "zra::DecompressBuffer(", "zra::DecompressRA(", "zra::Compressor::Compress(",
                "zra::Decompressor::Decompress(", "zra::FullDecompressor::Decompress(", "zra::Header::Header(",
                "zra::GetOutputBufferSize(", "zra::Exception::what("):
        assert sym in out, sym


def test_product_does_not_reference_the_oracle():
    import subprocess

    out = subprocess.run(["ldd", zra_b200.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "zra_ref" not in out
    src = os.path.join(ROOT, "zra_b200")
    for dirpath, _, files in os.walk(src):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".c", ".py")):
                text = open(os.path.join(dirpath, f)).read()
                assert "zra_oracle" not in text and "libzra_ref" not in text, f


def test_version_sizes_strings():
    assert zra_b200.GetVersion() == 1
    o = refzra.oracle()
    for n, fs in ((0, 16384), (1, 16384), (100_000, 16384), (1 << 20, 65536), (12345, 1000), (1 << 30, 65536)):
        assert zra_b200.GetOutputBufferSize(n, fs) == o.zra_oracle_output_buffer_size(n, fs, 0)
    L = zra_b200.lib()
    from zra_b200.binding import ZraStatus

    assert L.ZraGetErrorString(ZraStatus(5, 0)) == b"The specified offset and size are past the data contained within the buffer"
    assert L.ZraGetErrorString(ZraStatus(1, 20)) == b"An error was returned by ZStandard: Corrupted block detected"


def test_header_parsing_needs_no_gpu():
    archive, meta = golden_archive("text_f16384_l3")
    h = zra_b200.Header(archive)
    p = parse_header(archive)
    assert (h.version, h.size, h.uncompressedSize, h.frameSize, h.metaSize) == (1, p["size"], meta["bytes"], 16384, 0)
    # through a read callback
    h2 = zra_b200.Header(lambda off, size: archive[off: off + size].tobytes())
    assert h2.size == p["size"]
    bad = archive.copy()
    bad[8] ^= 1
    with pytest.raises(zra_b200.ZraError) as e:
        zra_b200.Header(bad)
    assert e.value.code == zra_b200.StatusCode.HeaderInvalid
    with pytest.raises(zra_b200.ZraError) as e:
        zra_b200.Header(archive[:38])  # the reference's `>=` bound: 38 bytes are not enough
    assert e.value.code == zra_b200.StatusCode.OutOfBoundsAccess


def test_argument_checks_precede_gpu_work():
    archive, meta = golden_archive("text_f16384_l3")
    n = meta["bytes"]
    with pytest.raises(zra_b200.ZraError) as e:
        zra_b200.DecompressRA(archive, n - 10, 10)  # last byte unreachable, like the reference
    assert e.value.code == zra_b200.StatusCode.OutOfBoundsAccess


def test_gpu_entry_points_fail_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    archive, _ = golden_archive("text_f16384_l3")
    with pytest.raises(zra_b200.ZraError) as e:
        zra_b200.DecompressBuffer(archive)
    assert e.value.code == zra_b200.StatusCode.ZStdError and e.value.zstd_code == 1
    with pytest.raises(zra_b200.ZraError):
        zra_b200.CudaContext(0)


def test_header_crc_verification_needs_no_gpu():
    """ZraVerifyHeaderCrc: the CRC-32 the reference writes (zra.cpp:128-133) and never checks. Every
    reference-made golden archive verifies; any flipped header byte (fixed fields, metadata, seek table) is caught."""
    import numpy as np
    import pytest

    import zra_b200
    from common import golden_archive, golden_archives, parse_header

    for name in golden_archives():
        a, _ = golden_archive(name)
        assert zra_b200.VerifyHeaderCrc(a), name
        h = parse_header(a)
        for at in (5, 19, 27, 31, 38 + h["metaSize"], h["size"] - 1, 15):
            if at >= h["size"]:
                continue
            b = a.copy()
            b[at] ^= 0x40
            assert not zra_b200.VerifyHeaderCrc(b), (name, at)
        with pytest.raises(zra_b200.ZraError):
            zra_b200.VerifyHeaderCrc(a[: h["size"] - 1])


def _put40(a, at, v):
    for i in range(5):
        a[at + i] = (v >> (8 * i)) & 255


def test_crafted_seek_tables_are_refused_before_any_gpu_work():
    """the seek table is untrusted. Entries out of order, outside the requested range or longer than 4 GiB
    must be refused with ZStdError / srcSize_wrong (72) BEFORE sizes are derived from them (these checks sit ahead of
    the first GPU call, so this runs in the CPU tier); nothing may unwind through the C shim."""
    archive, meta = golden_archive("text_f16384_l3")
    h = parse_header(archive)
    table = 38 + h["metaSize"]
    reader = lambda a: (lambda off, size: a[off: off + size].tobytes())
    fs = h["frameSize"]

    def expect_72(fn):
        with pytest.raises(zra_b200.ZraError) as e:
            fn()
        assert e.value.code == zra_b200.StatusCode.ZStdError and e.value.zstd_code == 72, (e.value.code, e.value.zstd_code)

    # entry 2 below entry 1 (fb < fa)
    bad = archive.copy()
    _put40(bad, table + 5 * 2, 1)
    expect_72(lambda: zra_b200.Decompressor(reader(bad)).Decompress(fs, 2 * fs))
    expect_72(lambda: zra_b200.FullDecompressor(reader(bad)).Decompress(np.empty(4 * fs, np.uint8)))
    # last entry of the range below the first (b < a): the compressed size would wrap
    bad = archive.copy()
    _put40(bad, table + 5 * 3, 0)
    expect_72(lambda: zra_b200.Decompressor(reader(bad)).Decompress(2 * fs, fs))
    # an inner entry beyond the end of the range
    bad = archive.copy()
    _put40(bad, table + 5 * 1, (1 << 39))
    expect_72(lambda: zra_b200.Decompressor(reader(bad)).Decompress(0, 2 * fs))


def test_crafted_header_geometry_is_refused_by_the_buffer_entry_points():
    """metaSize / tableSize must describe a seek table that lies inside the header (the frame-parallel decoder
    reads it; the reference's serial decoder never does): HeaderInvalid, not a wild host read."""
    archive, meta = golden_archive("text_f16384_l3")
    bad = archive.copy()
    bad[34:38] = np.frombuffer((0x7FFFFFF0).to_bytes(4, "little"), np.uint8)   # metaSize
    for fn in (lambda: zra_b200.DecompressBuffer(bad), lambda: zra_b200.DecompressRA(bad, 0, 100)):
        with pytest.raises(zra_b200.ZraError) as e:
            fn()
        assert e.value.code in (zra_b200.StatusCode.HeaderInvalid, zra_b200.StatusCode.OutOfBoundsAccess), e.value.code
