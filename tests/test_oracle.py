"""Pins the oracle (oracle/zra_oracle.c): known answers, committed golden vectors, and — where the
reference has been compiled (oracle/_ref) — the reference itself on fresh seeded inputs."""
import zlib

import numpy as np
import pytest

import refzra
from common import MANIFEST, golden_archive, golden_archives, golden_frame, golden_frames, parse_header, seek_table, sha
from zra_b200 import synth

needs_ref = pytest.mark.skipif(not refzra.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def test_crc32_known_answers():
    assert refzra.oracle_crc32(b"123456789") == 0xCBF43926  # CRCpp/test/src/main.cpp:255
    assert refzra.oracle_crc32(b"") == 0
    data = synth.random_bytes(10_000, seed=3, threads=1)
    assert refzra.oracle_crc32(data) == zlib.crc32(data.tobytes())
    a, b = data[:1234], data[1234:]
    assert refzra.oracle_crc32(b, refzra.oracle_crc32(a)) == zlib.crc32(data.tobytes())


def test_xxh64_known_answers():
    # XXH64 test vectors from the xxHash specification (seed 0)
    assert refzra.oracle_xxh64(b"") == 0xEF46DB3751D8E999
    assert refzra.oracle_xxh64(b"a") == 0xD24EC4F1A98C6E5B
    assert refzra.oracle_xxh64(b"abc") == 0x44BC2CF5AD770999
    assert refzra.oracle_xxh64(b"Nobody inspects the spammish repetition") == 0xFBCEA83C8A378BF1


@needs_ref
def test_xxh64_and_crc_match_reference():
    L = refzra.ref()
    for n in (0, 1, 3, 4, 7, 8, 31, 32, 33, 63, 64, 100, 4096, 65536 + 5):
        d = synth.random_bytes(n, seed=n + 1, threads=1)
        assert refzra.oracle_xxh64(d) == L.XXH64(refzra._p(d), n, 0)
        assert refzra.oracle_crc32(d) == L.ref_crc32(refzra._p(d), n, 0, 0)


def test_sizes():
    o = refzra.oracle()
    assert o.zra_oracle_output_buffer_size(100_000, 16384, 0) == 115_606  # SURVEY.md Z5 [probed]
    assert o.zra_oracle_compress_bound(0) == 64
    assert o.zra_oracle_compress_bound(128 << 10) == (128 << 10) + 512


@needs_ref
def test_sizes_match_reference():
    o, L = refzra.oracle(), refzra.ref()
    for s in (0, 1, 100, 4095, 16384, 65536, 131071, 131072, 131073, 1 << 20, (1 << 32) - 1):
        assert o.zra_oracle_compress_bound(s) == L.ZSTD_compressBound(s)
    for n, fs in ((0, 16384), (1, 16384), (100_000, 16384), (1 << 20, 65536), (12345, 1000)):
        assert o.zra_oracle_output_buffer_size(n, fs, 0) == L.ZraGetCompressedOutputBufferSize(n, fs)


@pytest.mark.parametrize("name", golden_archives())
def test_golden_archives_decode(name):
    archive, meta = golden_archive(name)
    out = refzra.oracle_decompress_buffer(archive)
    assert out.size == meta["bytes"]
    assert sha(out) == meta["sha256"]
    # header: layout constants and CRC (the reference writes but never verifies it)
    h = parse_header(archive)
    assert h["frameId"] == 0x184D2A50 and h["magic"] == 0x3041525A and h["version"] == 1
    assert h["frameSize"] == meta["frameSize"] and h["uncompressedSize"] == meta["bytes"]
    assert refzra.oracle().zra_oracle_header_crc(refzra._p(archive), archive.size) == h["hash"]
    t = seek_table(archive)
    assert h["size"] + int(t[-1]) == archive.size
    # stock zstd skips the header frame and decodes the rest
    assert sha(refzra.system_zstd_decompress(archive, meta["bytes"])) == meta["sha256"]


@pytest.mark.parametrize("name", golden_archives())
def test_golden_header_rebuild(name):
    """zra_oracle_build_header regenerates the reference's header bytes from the seek table."""
    import ctypes as C

    archive, meta = golden_archive(name)
    h = parse_header(archive)
    t = seek_table(archive)
    offs = (C.c_uint64 * len(t))(*[int(x) for x in t])
    out = np.zeros(h["size"], np.uint8)
    n = refzra.oracle().zra_oracle_build_header(refzra._p(out), h["uncompressedSize"], h["frameSize"], None, 0, offs, len(t))
    assert n == h["size"]
    assert np.array_equal(out, archive[: h["size"]])


def test_decodecorpus_frames():
    bad = []
    for name in golden_frames():
        z, meta = golden_frame(name)
        try:
            out = refzra.oracle_zstd_decompress(z, meta["bytes"])
        except refzra.OracleError as e:
            bad.append((name, str(e)))
            continue
        if out.size != meta["bytes"] or sha(out) != meta["sha256"]:
            bad.append((name, "mismatch"))
    assert not bad, bad


def test_zstd_golden_decompression_file():
    import os

    from common import GOLDEN

    z = np.fromfile(os.path.join(GOLDEN, "rle-first-block.zst"), dtype=np.uint8)
    out = refzra.oracle_zstd_decompress(z, 2 << 20)
    assert out.size == 1 << 20 and not out.any()  # 1 MiB of zeros: an RLE first block
    assert np.array_equal(out, refzra.system_zstd_decompress(z, 2 << 20))


def test_random_access_semantics():
    archive, meta = golden_archive("text_f16384_l3")
    full = refzra.oracle_decompress_buffer(archive)
    n = meta["bytes"]
    for off, size in ((0, 1), (0, 16384), (1, 16384), (16383, 2), (16384, 16384), (5000, 70000), (n - 10, 9), (123, 0)):
        assert np.array_equal(refzra.oracle_decompress_ra(archive, off, size), full[off: off + size])
    # the in-memory entry point cannot reach the last byte (`>=`), the streaming one can (`>`)
    with pytest.raises(refzra.OracleError) as e:
        refzra.oracle_decompress_ra(archive, n - 10, 10)
    assert e.value.zra == 5
    assert np.array_equal(refzra.oracle_decompress_ra(archive, n - 10, 10, in_memory_quirk=False), full[n - 10:])


def test_corruption_is_detected():
    archive, meta = golden_archive("text_f16384_l3")
    h = parse_header(archive)
    bad = archive.copy()
    bad[h["size"] + 40] ^= 0x55
    with pytest.raises(refzra.OracleError) as e:
        refzra.oracle_decompress_buffer(bad)
    assert e.value.zra == 1 and e.value.zstd in (20, 22)
    bad = archive.copy()
    bad[8] ^= 1
    with pytest.raises(refzra.OracleError) as e:
        refzra.oracle_decompress_buffer(bad)
    assert e.value.zra == 3


@needs_ref
@pytest.mark.parametrize("kind,fs,lvl,ck", [("text", 16384, 3, True), ("text", 65536, 1, False), ("text", 262144, 3, True),
                                            ("mixed", 65536, 2, True), ("text", 5000, 7, True), ("random", 16384, 3, True),
                                            ("datagen30", 16384, 3, True), ("datagen70", 262144, 1, True), ("datagen95", 65536, 9, True)])
def test_fresh_reference_archives(kind, fs, lvl, ck):
    n = 1_000_003
    if kind.startswith("datagen"):  # zstd's own generator (programs/datagen.c): other match / literal statistics than the text model
        import os
        import subprocess

        gen = os.path.join(os.path.dirname(refzra.REF_SO), "datagen")
        if not os.path.exists(gen):
            pytest.skip("oracle/_ref/datagen not present")
        data = np.frombuffer(subprocess.run([gen, f"-g{n}", f"-P{kind[7:]}", "-s3"], capture_output=True, check=True).stdout,
                             dtype=np.uint8)[:n].copy()
    else:
        data = {"text": synth.text, "random": synth.random_bytes}.get(kind, None)
        data = synth.mixed(n, period=fs, threads=1) if kind == "mixed" else data(n, seed=fs + lvl, threads=1)
    z = refzra.ref_compress(data, lvl, fs, ck)
    assert np.array_equal(refzra.oracle_decompress_buffer(z), data)
    assert np.array_equal(refzra.ref_decompress(z), data)
    rng = np.random.default_rng(5)
    for _ in range(20):
        off = int(rng.integers(0, n - 1))
        size = int(rng.integers(0, min(3 * fs, n - off - 1)))
        assert np.array_equal(refzra.oracle_decompress_ra(z, off, size), refzra.ref_decompress_ra(z, off, size))


@needs_ref
def test_harness_stitching_equals_serial_reference():
    """ref_compress_mt (T Compressors + stitched tables) is byte-identical to zra::CompressBuffer."""
    data = synth.text(700_001, seed=9, threads=1)
    for fs, lvl in ((16384, 3), (65536, 1)):
        a = refzra.ref_compress(data, lvl, fs, True)
        b = refzra.ref_compress_mt(data, lvl, fs, True, threads=5)
        assert np.array_equal(a, b)


@needs_ref
def test_error_codes_match_reference():
    archive, _ = golden_archive("text_f65536_l3")
    h = parse_header(archive)
    rng = np.random.default_rng(11)
    agree = 0
    for _ in range(60):
        bad = archive.copy()
        pos = int(rng.integers(h["size"], archive.size))
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            refzra.ref_decompress(bad)
            ref_err = None
        except refzra.OracleError as e:
            ref_err = (e.zra, e.zstd)
        try:
            refzra.oracle_decompress_buffer(bad)
            ora_err = None
        except refzra.OracleError as e:
            ora_err = (e.zra, e.zstd)
        # both must agree on success/failure; the zstd sub-code may differ for exotic corruptions
        assert (ref_err is None) == (ora_err is None)
        agree += ref_err == ora_err
    assert agree >= 50
