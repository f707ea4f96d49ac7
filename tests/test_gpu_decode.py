"""GPU parity tests of the decode path, all through the C-ABI of libzra_b200.so.

Checker = the oracle (oracle/libzra_oracle.so), the committed golden vectors, and the unmodified
reference where its prebuilt .so travelled with the snapshot (oracle/_ref)."""
import numpy as np
import pytest

import refzra
import zra_b200
from common import golden_archive, golden_archives, golden_frame, golden_frames, parse_header, seek_table, sha
from zra_b200 import synth

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not refzra.have_ref(), reason="oracle/_ref not present")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(scope="module")
def ctx(torch_cuda):
    return zra_b200.CudaContext(0)


def to_device(torch, a, slack=16):
    t = torch.zeros(a.size + slack, dtype=torch.uint8, device="cuda")
    if a.size:
        t[: a.size] = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t


# ------------------------------------------------------------------ frame-range shards (what each rank of a box runs)
@pytest.mark.parametrize("name,shards", [("text_f16384_l3", 3), ("text_f262144_l3", 2), ("mixed_f16384_l3", 8), ("onebyte_f16384_l3", 2),
                                         ("text_f1000_l5", 5)])
def test_frame_range_shards_concatenate_to_the_whole(torch_cuda, ctx, name, shards):
    """ZraCudaDecompressFrames over contiguous frame ranges [g*F/G, (g+1)*F/G) regenerates, shard by shard,
    exactly the bytes of the whole archive; every shard writes only its own range."""
    from zra_b200 import shard

    torch = torch_cuda
    archive, meta = golden_archive(name)
    h = parse_header(archive)
    frames = h["tableSize"] - 1
    fs = h["frameSize"]
    d_in = to_device(torch, archive)
    got = np.zeros(meta["bytes"], np.uint8)
    for g in range(shards):
        f0, f1 = shard.frame_range(frames, g, shards)
        lo, hi = min(f0 * fs, meta["bytes"]), min(f1 * fs, meta["bytes"])
        d_out = torch.full((hi - lo + 64,), 0x55, dtype=torch.uint8, device="cuda")
        ctx.decompress_frames(d_in.data_ptr(), archive.size, f0, f1 - f0, d_out.data_ptr(), hi - lo, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        host = d_out.cpu().numpy()
        assert (host[hi - lo:] == 0x55).all()
        got[lo:hi] = host[: hi - lo]
    assert sha(got) == meta["sha256"]


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", golden_archives())
def test_golden_archives_host_api(name):
    archive, meta = golden_archive(name)
    out = zra_b200.DecompressBuffer(archive)
    assert out.size == meta["bytes"] and sha(out) == meta["sha256"]


@pytest.mark.parametrize("name", golden_archives())
def test_golden_archives_device_api(torch_cuda, ctx, name):
    torch = torch_cuda
    archive, meta = golden_archive(name)
    d_in = to_device(torch, archive)
    d_out = torch.full((meta["bytes"] + 64,), 0xAA, dtype=torch.uint8, device="cuda")
    ctx.decompress_buffer(d_in.data_ptr(), archive.size, d_out.data_ptr(), meta["bytes"], torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = d_out.cpu().numpy()
    assert sha(host[: meta["bytes"]]) == meta["sha256"]
    assert (host[meta["bytes"]:] == 0xAA).all()  # nothing written past the end


def test_decodecorpus_frames(torch_cuda, ctx):
    """All generative conformance frames in ONE batch (frames of very different shapes side by side)."""
    torch = torch_cuda
    blobs, frames, expect = [], [], []
    src_off = dst_off = 0
    for name in golden_frames():
        z, meta = golden_frame(name)
        pad = (-z.size) % 4
        blobs.append(np.concatenate([z, np.zeros(pad, np.uint8)]))
        frames.append((src_off, z.size, dst_off, meta["bytes"], 0))
        expect.append((dst_off, meta))
        src_off += z.size + pad
        dst_off += meta["bytes"] + 3
    src = np.concatenate(blobs)
    d_in = to_device(torch, src)
    d_out = torch.zeros(dst_off + 64, dtype=torch.uint8, device="cuda")
    sizes = ctx.decode_frames(d_in.data_ptr(), src.size, frames, d_out.data_ptr(), want_sizes=True)
    host = d_out.cpu().numpy()
    bad = [m["name"] for (off, m), sz in zip(expect, sizes) if sz != m["bytes"] or sha(host[off: off + m["bytes"]]) != m["sha256"]]
    assert not bad, bad


def test_zstd_golden_rle_first_block(torch_cuda, ctx):
    import os

    from common import GOLDEN

    torch = torch_cuda
    z = np.fromfile(os.path.join(GOLDEN, "rle-first-block.zst"), dtype=np.uint8)
    d_in = to_device(torch, z)
    d_out = torch.full(((1 << 20) + 16,), 7, dtype=torch.uint8, device="cuda")
    sizes = ctx.decode_frames(d_in.data_ptr(), z.size, [(0, z.size, 0, 1 << 20, 0)], d_out.data_ptr(), want_sizes=True)
    assert sizes == [1 << 20]
    assert not d_out[: 1 << 20].any().item()


# ------------------------------------------------------------------ random access
@pytest.mark.parametrize("name", ["text_f16384_l3", "text_f262144_l3", "text_f1000_l5", "mixed_f16384_l3"])
def test_random_access_in_memory(name):
    archive, meta = golden_archive(name)
    full = refzra.oracle_decompress_buffer(archive)
    n, fs = meta["bytes"], meta["frameSize"]
    rng = np.random.default_rng(1)
    cases = [(0, 1), (0, fs), (1, fs), (fs - 1, 2), (fs, fs), (n - 10, 9), (123, 0), (fs // 2, 3 * fs)]
    cases += [(int(o), int(s)) for o, s in zip(rng.integers(0, n - 1, 25), rng.integers(0, 3 * fs, 25))]
    for off, size in cases:
        size = min(size, n - off - 1)
        got = zra_b200.DecompressRA(archive, off, size)
        assert np.array_equal(got, refzra.oracle_decompress_ra(archive, off, size)), (off, size)
        assert np.array_equal(got, full[off: off + size])
    with pytest.raises(zra_b200.ZraError) as e:
        zra_b200.DecompressRA(archive, n - 10, 10)
    assert e.value.code == zra_b200.StatusCode.OutOfBoundsAccess


@pytest.mark.parametrize("name", ["text_f16384_l3", "text_f262144_l3", "text_f1000_l5", "mixed_f16384_l3"])
def test_random_access_batch(torch_cuda, ctx, name):
    """ZraCudaDecompressRABatch: every read equals the oracle's DecompressRA / a slice of the full output; frames are
    de-duplicated across the batch."""
    torch = torch_cuda
    archive, meta = golden_archive(name)
    full = refzra.oracle_decompress_buffer(archive)
    n, fs = meta["bytes"], meta["frameSize"]
    d_in = to_device(torch, archive)
    rng = np.random.default_rng(5)
    # uniform 4 KiB reads (the BASELINE shape), heavy frame sharing
    size = min(4096, n // 2)
    offs = np.concatenate([rng.integers(0, n - size + 1, 700), [0, n - size, fs - 1, max(0, fs - size)]]).astype(np.uint64)
    d_off = torch.from_numpy(offs.view(np.int64)).cuda()
    d_out = torch.full((offs.size * size,), 0xAA, dtype=torch.uint8, device="cuda")
    unique = ctx.decompress_ra_batch(d_in.data_ptr(), archive.size, d_off.data_ptr(), offs.size, d_out.data_ptr(), uniform_size=size)
    got = d_out.cpu().numpy().reshape(offs.size, size)
    touched = set()
    for i, o in enumerate(offs):
        o = int(o)
        assert np.array_equal(got[i], full[o: o + size]), (i, o)
        touched.update(range(o // fs, (o + size - 1) // fs + 1))
    assert unique == len(touched)
    for i in (0, 5, 11):  # against the oracle's restatement of the reference's streaming random access
        assert np.array_equal(got[i], refzra.oracle_decompress_ra(archive, int(offs[i]), size, in_memory_quirk=False))
    # a second batch on the same context (the frame->slot map must have been reset)
    unique2 = ctx.decompress_ra_batch(d_in.data_ptr(), archive.size, d_off.data_ptr(), 10, d_out.data_ptr(), uniform_size=size)
    assert unique2 == len({f for o in offs[:10] for f in range(int(o) // fs, (int(o) + size - 1) // fs + 1)})
    # ragged reads with explicit output offsets, including empty reads and the very last byte
    sizes = rng.integers(0, min(3 * fs, n), 200).astype(np.uint32)
    roffs = np.array([rng.integers(0, n - int(s) + 1) for s in sizes], dtype=np.uint64)
    sizes[0], roffs[0] = 1, n - 1
    outo = np.concatenate([[0], np.cumsum(sizes.astype(np.uint64))[:-1]]).astype(np.uint64)
    total = int(sizes.sum())
    d_o = torch.from_numpy(roffs.view(np.int64)).cuda()
    d_s = torch.from_numpy(sizes.view(np.int32)).cuda()
    d_oo = torch.from_numpy(outo.view(np.int64)).cuda()
    d_out2 = torch.zeros(total + 16, dtype=torch.uint8, device="cuda")
    ctx.decompress_ra_batch(d_in.data_ptr(), archive.size, d_o.data_ptr(), sizes.size, d_out2.data_ptr(), d_sizes=d_s.data_ptr(),
                            d_out_offsets=d_oo.data_ptr(), max_size=int(sizes.max()))
    got2 = d_out2.cpu().numpy()
    for o, s, w in zip(roffs, sizes, outo):
        assert np.array_equal(got2[int(w): int(w) + int(s)], full[int(o): int(o) + int(s)])
    # bounds: offset + size > uncompressedSize is OutOfBoundsAccess (zra::Decompressor::Decompress), with the request index
    bad = offs.copy()
    bad[7] = n - size + 1
    d_bad = torch.from_numpy(bad.view(np.int64)).cuda()
    with pytest.raises(zra_b200.ZraError) as e:
        ctx.decompress_ra_batch(d_in.data_ptr(), archive.size, d_bad.data_ptr(), bad.size, d_out.data_ptr(), uniform_size=size)
    assert e.value.code == zra_b200.StatusCode.OutOfBoundsAccess and e.value.bad_request == 7
    # ... and the context still works afterwards
    assert ctx.decompress_ra_batch(d_in.data_ptr(), archive.size, d_off.data_ptr(), 10, d_out.data_ptr(), uniform_size=size) == unique2


def test_streaming_decompressor_and_full_decompressor():
    archive, meta = golden_archive("text_f16384_l3")
    full = refzra.oracle_decompress_buffer(archive)
    n, fs = meta["bytes"], meta["frameSize"]
    calls = []

    def reader(off, size):
        calls.append((off, size))
        return archive[off: off + size].tobytes()

    d = zra_b200.Decompressor(reader)
    assert d.header.uncompressedSize == n
    rng = np.random.default_rng(2)
    for off, size in [(0, n), (n - 10, 10), (5, 0), (fs, fs)] + [(int(o), 4096) for o in rng.integers(0, n - 4096, 20)]:
        calls.clear()
        got = d.Decompress(off, size)
        assert np.array_equal(got, full[off: off + size]), (off, size)
        assert len(calls) == 1  # exactly one read callback per request, like the reference
        assert np.array_equal(got, refzra.oracle_decompress_ra(archive, off, size, in_memory_quirk=False))
    with pytest.raises(zra_b200.ZraError):
        d.Decompress(n - 5, 6)

    fd = zra_b200.FullDecompressor(reader)
    out = np.empty(3 * fs + 100, np.uint8)  # room for 3 frames per call
    pieces = []
    while True:
        k = fd.Decompress(out)
        if not k:
            break
        assert k <= 3 * fs
        pieces.append(out[:k].copy())
    assert np.array_equal(np.concatenate(pieces), full)
    small = np.empty(fs - 1, np.uint8)
    with pytest.raises(zra_b200.ZraError) as e:
        zra_b200.FullDecompressor(reader).Decompress(small)
    assert e.value.code == zra_b200.StatusCode.OutputBufferTooSmall


# ------------------------------------------------------------------ errors
def test_corruption_matches_oracle():
    archive, _ = golden_archive("text_f16384_l3")
    h = parse_header(archive)
    rng = np.random.default_rng(3)
    exact = 0
    for i in range(40):
        bad = archive.copy()
        pos = int(rng.integers(h["size"], archive.size))
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            refzra.oracle_decompress_buffer(bad)
            ora = None
        except refzra.OracleError as e:
            ora = (e.zra, e.zstd)
        try:
            zra_b200.DecompressBuffer(bad)
            got = None
        except zra_b200.ZraError as e:
            got = (int(e.code), e.zstd_code)
        assert (got is None) == (ora is None), (pos, got, ora)
        exact += got == ora
    assert exact >= 38   # the known divergence: a literals section declared larger than the frame is refused before its Huffman streams are read (70 where zstd may report 20)
    bad = archive.copy()
    bad[8] ^= 1
    with pytest.raises(zra_b200.ZraError) as e:
        zra_b200.DecompressBuffer(bad)
    assert e.value.code == zra_b200.StatusCode.HeaderInvalid
    with pytest.raises(zra_b200.ZraError) as e:
        zra_b200.DecompressBuffer(archive[:-7])  # truncated
    assert e.value.code == zra_b200.StatusCode.ZStdError and e.value.zstd_code == 72


# ------------------------------------------------------------------ bigger, fresh archives
@needs_ref
@pytest.mark.parametrize("kind,fs,lvl,ck,n", [
    ("text", 16384, 3, True, 32 << 20),
    ("text", 65536, 3, True, 64 << 20),
    ("text", 65536, 1, False, 32 << 20),
    ("text", 262144, 3, True, (32 << 20) + 12345),
    ("mixed", 65536, 2, True, 32 << 20),
    ("text", 1 << 20, 5, True, 16 << 20),
    ("text", 4 << 20, 1, True, (24 << 20) + 999),   # 32 blocks per frame: repeat-mode tables, treeless literals across rounds
    ("mixed", 1 << 20, 3, False, 12 << 20),
    ("text", 3000, 9, True, 3 << 20),
    ("zeros", 65536, 3, True, 16 << 20),
    ("datagen40", 65536, 3, True, 24 << 20),    # zstd's own generator: other match / literal statistics than the text model
    ("datagen90", 16384, 5, True, 8 << 20),
    ("datagen10", 262144, 1, True, 16 << 20),
])
def test_reference_archives_roundtrip(torch_cuda, ctx, kind, fs, lvl, ck, n):
    """encode (reference, host cores) -> decode (CUDA): size-independent round-trip property."""
    torch = torch_cuda
    if kind.startswith("datagen"):
        import os
        import subprocess

        gen = os.path.join(os.path.dirname(refzra.REF_SO), "datagen")
        if not os.path.exists(gen):
            pytest.skip("oracle/_ref/datagen not present")
        data = np.frombuffer(subprocess.run([gen, f"-g{n}", f"-P{kind[7:]}", "-s5"], capture_output=True, check=True).stdout,
                             dtype=np.uint8)[:n].copy()
    else:
        data = (synth.mixed(n, period=fs) if kind == "mixed" else np.zeros(n, np.uint8) if kind == "zeros" else synth.text(n, seed=fs + lvl))
    z = refzra.ref_compress_mt(data, lvl, fs, ck)
    got = zra_b200.DecompressBuffer(z)
    assert np.array_equal(got, data)
    # device-resident path + frame-range form
    d_in = to_device(torch, z)
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.decompress_buffer(d_in.data_ptr(), z.size, d_out.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    assert torch.equal(d_out, torch.from_numpy(data).cuda())
    frames = (n + fs - 1) // fs
    f0, cnt = frames // 3, max(1, frames // 2)
    cnt = min(cnt, frames - f0)
    lo, hi = f0 * fs, min(n, (f0 + cnt) * fs)
    d_part = torch.zeros(hi - lo, dtype=torch.uint8, device="cuda")
    ctx.decompress_frames(d_in.data_ptr(), z.size, f0, cnt, d_part.data_ptr(), hi - lo, torch.cuda.current_stream().cuda_stream)
    assert torch.equal(d_part, torch.from_numpy(data[lo:hi]).cuda())


# ------------------------------------------------------------------ scratch budget: archives larger than one group
@needs_ref
@pytest.mark.parametrize("fs,env", [(16384, {"ZRA_B200_SCRATCH_MB": "64", "ZRA_B200_CHUNKS": "3"}),
                                    (262144, {"ZRA_B200_SCRATCH_MB": "64", "ZRA_B200_CHUNKS": "7"}),
                                    (65536, {"ZRA_B200_CHUNKS": "40"})])
def test_group_and_chunk_cuts_do_not_change_the_result(tmp_path, fs, env):
    """BASELINE's full sizes (8 GiB shards) do not fit one scratch group: the decode then runs group after group, each
    cut into chunks. Forced here at a small size through the tuning environment (read once per process, hence the
    subprocess): device path, host path and batched random access must give the reference's bytes."""
    import os
    import subprocess
    import sys

    n = 48 << 20
    data = synth.text(n, seed=fs)
    archive = refzra.ref_compress_mt(data, 3, fs, True)
    np.asarray(archive).tofile(tmp_path / "a.zra")
    data.tofile(tmp_path / "a.bin")
    code = f"""
import sys, numpy as np, torch
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import zra_b200
a = np.fromfile({str(tmp_path / 'a.zra')!r}, dtype=np.uint8); d = np.fromfile({str(tmp_path / 'a.bin')!r}, dtype=np.uint8)
assert np.array_equal(zra_b200.DecompressBuffer(a), d), "host path"
ctx = zra_b200.CudaContext(0)
d_in = torch.zeros(a.size + 64, dtype=torch.uint8, device="cuda"); d_in[:a.size] = torch.from_numpy(a).cuda()
d_out = torch.empty(d.size, dtype=torch.uint8, device="cuda")
ctx.decompress_buffer(d_in.data_ptr(), a.size, d_out.data_ptr(), d.size, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
assert np.array_equal(d_out.cpu().numpy(), d), "device path"
rng = np.random.default_rng(5); offs = rng.integers(0, d.size - 4097, 4096).astype(np.uint64)
d_off = torch.from_numpy(offs.view(np.int64)).cuda(); d_ra = torch.empty(4096 * 4096, dtype=torch.uint8, device="cuda")
ctx.decompress_ra_batch(d_in.data_ptr(), a.size, d_off.data_ptr(), 4096, d_ra.data_ptr(), uniform_size=4096, stream=torch.cuda.current_stream().cuda_stream)
got = d_ra.cpu().numpy().reshape(4096, 4096)
for i in range(0, 4096, 37): assert np.array_equal(got[i], d[int(offs[i]): int(offs[i]) + 4096]), "random access"
print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-1500:]


def test_full_decompressor_read_ahead(torch_cuda):
    """FullDecompressor reads the NEXT call's bytes through the user's callback while the GPU works on this call
    (SURVEY.md 8f-1). Same total number of callbacks and the same ranges as the reference's callback-per-call loop, each
    range read exactly once when the caller keeps its buffer size; a caller that changes it, and two decompressors
    interleaved on one thread (they share the thread's staging buffers), still get the right bytes."""
    archive, meta = golden_archive("text_f16384_l3")
    full = refzra.oracle_decompress_buffer(archive)
    fs = meta["frameSize"]
    calls = []

    def reader(off, size):
        calls.append((off, size))
        return archive[off: off + size].tobytes()

    def drain(fd, sizes):
        pieces, i = [], 0
        while True:
            out = np.empty(sizes[i % len(sizes)], np.uint8)
            i += 1
            k = fd.Decompress(out)
            if not k:
                return np.concatenate(pieces)
            pieces.append(out[:k].copy())

    # steady buffer size: every compressed byte is read exactly once, in order
    calls.clear()
    fd = zra_b200.FullDecompressor(reader)
    hdr_calls = len(calls)
    assert np.array_equal(drain(fd, [3 * fs]), full)
    data_calls = [c for c in calls[hdr_calls:] if c[1]]
    assert [c[0] for c in data_calls] == sorted(c[0] for c in data_calls)
    assert sum(c[1] for c in data_calls) == archive.size - parse_header(archive)["size"]
    # a caller that changes its buffer size between calls: a read-ahead that does not fit is dropped, bytes stay right
    assert np.array_equal(drain(zra_b200.FullDecompressor(reader), [3 * fs, fs, 5 * fs + 7, 2 * fs]), full)
    # two decompressors interleaved on one thread
    a, b = zra_b200.FullDecompressor(reader), zra_b200.FullDecompressor(reader)
    pa, pb = [], []
    oa, ob = np.empty(2 * fs, np.uint8), np.empty(3 * fs, np.uint8)
    done_a = done_b = False
    while not (done_a and done_b):
        if not done_a:
            k = a.Decompress(oa)
            done_a = k == 0
            pa.append(oa[:k].copy())
        if not done_b:
            k = b.Decompress(ob)
            done_b = k == 0
            pb.append(ob[:k].copy())
    assert np.array_equal(np.concatenate(pa), full) and np.array_equal(np.concatenate(pb), full)
