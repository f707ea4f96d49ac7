"""Multi-GPU sharding logic on CPU: world_size-2 and -3 gloo process groups.

The exchange step of sharded compression — all-gather of the per-frame compressed sizes / exclusive
scan of the per-shard totals — and the header stitching (ZraShardBuildHeader, host-only C-ABI) are
checked against an archive the unmodified reference wrote in one piece: frames are independent
(SURVEY.md E11 [probed]), so the stitched archive must be byte-identical."""
import os
import socket
import sys

import numpy as np
import pytest

import refzra
from zra_b200 import shard, synth

needs_ref = pytest.mark.skipif(not refzra.have_ref(), reason="oracle/_ref not present")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, fs, level, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = synth.text(n, seed=5, threads=1)
        frames = (n + fs - 1) // fs
        lo, hi = shard.byte_range(n, fs, rank, world)
        f0, f1 = shard.frame_range(frames, rank, world)
        # this rank's frames, made by the reference (the CPU tier has no GPU encoder): a shard-local archive
        local = refzra.ref_compress(data[lo:hi], level, fs, True) if hi > lo else None
        if local is not None:
            hdr = 38 + 5 * (f1 - f0 + 1)
            t = local[38:hdr].reshape(-1, 5).astype(np.uint64)
            offs = t[:, 0] | (t[:, 1] << 8) | (t[:, 2] << 16) | (t[:, 3] << 24) | (t[:, 4] << 32)
            sizes, payload = np.diff(offs), local[hdr:]
        else:
            sizes, payload = np.zeros(0, np.uint64), np.zeros(0, np.uint8)
        all_sizes, base, total = shard.exchange_frame_sizes(sizes, frames)
        header = shard.build_header(n, fs, all_sizes)
        q.put((rank, base, total, header.tobytes(), payload.tobytes()))
    finally:
        dist.destroy_process_group()


@needs_ref
@pytest.mark.parametrize("world,n,fs", [(2, 1 << 20, 16384), (3, (1 << 20) + 777, 65536), (2, 40000, 16384)])
def test_sharded_compress_stitches_to_the_reference_archive(world, n, fs):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, fs, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = refzra.ref_compress(synth.text(n, seed=5, threads=1), 3, fs, True)
    header = results[0][3]
    assert all(r[3] == header for r in results)  # every rank stitched the same header
    total = results[0][2]
    body = bytearray(total)
    for rank, base, tot, _, payload in results:
        assert tot == total
        body[base: base + len(payload)] = payload  # the final gather: shard bytes at their scanned base offsets
    assert header + bytes(body) == whole.tobytes()


def test_frame_ranges_partition_the_archive():
    for frames in (0, 1, 7, 8, 1000, 16385):
        for world in (1, 2, 3, 8):
            rs = [shard.frame_range(frames, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == frames
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1


def test_build_header_matches_the_oracle():
    rng = np.random.default_rng(0)
    for frames, fs, meta in ((1, 16384, b""), (5, 1000, b"meta!"), (300, 65536, b"")):
        sizes = rng.integers(9, 70000, frames).astype(np.uint64)
        n = (frames - 1) * fs + 17
        got = shard.build_header(n, fs, sizes, meta)
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
        import ctypes as C

        o = refzra.oracle()
        out = np.zeros(38 + len(meta) + 5 * (frames + 1), np.uint8)
        m = np.frombuffer(meta, np.uint8)
        o.zra_oracle_build_header(out.ctypes.data_as(C.c_void_p), n, fs, m.ctypes.data_as(C.c_void_p) if m.size else None, m.size,
                                  offs.ctypes.data_as(C.POINTER(C.c_uint64)), frames + 1)
        assert np.array_equal(got, out)


# ---------------------------------------------------------------------------- random access over a sharded archive
def test_route_reads_splits_at_shard_boundaries():
    n, fs, world = 10 * 16384 + 100, 16384, 3
    offs = np.array([0, 16384 * 3 - 10, n - 5, 16384 * 6 - 1, 50000], dtype=np.int64)
    szs = np.array([100, 20, 5, 2 * 16384 + 2, 0], dtype=np.int64)
    routed = shard.route_reads(offs, szs, n, fs, world)
    seen = np.zeros(offs.size, np.int64)
    for r, (req, inner, absolute, size) in enumerate(routed):
        lo, hi = shard.byte_range(n, fs, r, world)
        assert ((absolute >= lo) & (absolute + size <= hi)).all()
        assert (absolute == offs[req] + inner).all()
        np.add.at(seen, req, size)
    assert (seen == szs).all()          # every byte of every read is served exactly once
    assert len(routed[0][0]) == 2       # read 1 straddles shards 0|1 (frames 0-2 | 3-5)
    with pytest.raises(Exception):
        shard.route_reads(np.array([n - 4]), np.array([5]), n, fs, world)


def _ra_worker(rank, world, port, n, fs, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = synth.text(n, seed=9, threads=1)
        lo, hi = shard.byte_range(n, fs, rank, world)
        mine = data[lo:hi]           # what this rank's GPU would hold decoded on demand

        def serve(abs_off, sizes):
            assert all(lo <= o and o + s <= hi for o, s in zip(abs_off, sizes))
            return b"".join(mine[o - lo: o - lo + s].tobytes() for o, s in zip(abs_off, sizes))

        rng = np.random.default_rng(100 + rank)
        cnt = 300
        sizes = rng.integers(0, 3 * fs, cnt).astype(np.int64)
        offs = np.array([rng.integers(0, n - s + 1) for s in sizes], dtype=np.int64)
        got = shard.sharded_random_access(offs, sizes, n, fs, serve)
        want = np.concatenate([data[o: o + s] for o, s in zip(offs, sizes)]) if cnt else np.zeros(0, np.uint8)
        q.put((rank, bool(np.array_equal(got, want)), int(got.size)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_random_access_over_gloo(world):
    """Every rank issues its own reads into the whole archive; pieces are routed to the owners of their frames, served
    from the owners' shards and reassembled in request order (reads crossing shard boundaries included)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n, fs = (1 << 20) + 4321, 16384
    procs = [ctx.Process(target=_ra_worker, args=(r, world, port, n, fs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res), res


def _reads_worker(rank, world, port, n, shard_bytes, rsz, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = synth.text(n, seed=9, threads=1)
        lo, hi = rank * shard_bytes, min(n, (rank + 1) * shard_bytes)
        mine = torch.from_numpy(data[lo:hi].copy())
        served = [0]

        def serve(abs_off, sizes):
            a, s = abs_off.tolist(), sizes.tolist()
            assert all(lo <= o and o + z <= hi for o, z in zip(a, s)), "a piece outside this rank's shard"
            served[0] += len(a)
            return torch.cat([mine[o - lo: o - lo + z] for o, z in zip(a, s)]) if a else torch.empty(0, dtype=torch.uint8)

        rng = np.random.default_rng(100 + rank)
        cnt = 500
        offs = rng.integers(0, n - rsz + 1, cnt).astype(np.int64)
        # make sure some reads straddle every shard boundary
        for b in range(1, world):
            offs[b] = b * shard_bytes - 1 - rank
            offs[world + b] = b * shard_bytes - rsz + 1
        sr = shard.ShardedReads(n, shard_bytes, rsz)
        got = sr.read(torch.from_numpy(offs), serve).numpy()
        want = np.stack([data[o: o + rsz] for o in offs])
        q.put((rank, bool(np.array_equal(got, want)), served[0]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_reads_all_to_all_over_gloo(world):
    """The tensor data plane bench.py uses on NCCL (ShardedReads: all_to_all_single of offsets, then of bytes), on gloo:
    every rank's reads into the whole archive come back bit-exact and in request order, boundary-crossing reads included."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    fs, rsz = 16384, 4096
    shard_bytes = 20 * fs
    n = world * shard_bytes - 1234     # the last shard is shorter
    procs = [ctx.Process(target=_reads_worker, args=(r, world, port, n, shard_bytes, rsz, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res), res
    assert sum(r[2] for r in res) >= world * 500     # every read was served by somebody (crossing reads twice)
