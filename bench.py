#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for zra-b200.

Workload (BASELINE.json configs[1]): DecompressBuffer of a 1 GiB ZRA archive, 64 KiB frames,
Zipf-word text, written by the UNMODIFIED reference at its default level 3 with checksums.
A "step" is one whole-archive decompression. Numbers on the JSON line:
  value     decompressed GB/s (10^9 bytes of ORIGINAL data per second), archive and output resident in
            HBM, through the device-pointer C-ABI (ZraCudaDecompressBuffer); CUDA events, max over ranks
  e2e       same metric through the reference-facing C-ABI ZraDecompressBuffer with pinned HOST buffers
            (H2D of the archive and D2H of the output inside the timed region)
  roofline  dominant kernel: algorithmic bytes (archive bytes read + original bytes written) of one
            step / that kernel's CUDA-event time per step, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the reference's own CPU path (oracle/_ref) on this box's host cores, same archive
`--impl reference` times the reference CPU implementation alone (all host threads) on the same archive.
N > 1 (torchrun): every rank decodes its own 1 GiB archive (weak scaling, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GIB = 1 << 30


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="zra_b200", choices=["zra_b200", "reference"])
    ap.add_argument("--size-mib", type=int, default=1024, help="original bytes per archive (default: the 1 GiB of configs[1])")
    ap.add_argument("--frame-size", type=int, default=65536)
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ra", action="store_true", help="skip the batched random-access leg")
    ap.add_argument("--ra-size-mib", type=int, default=1024, help="original bytes of the random-access archive (16 KiB frames)")
    ap.add_argument("--no-compress", action="store_true", help="skip the CompressBuffer leg")
    ap.add_argument("--no-streaming", action="store_true", help="skip the FullDecompressor streaming leg")
    ap.add_argument("--stream-size-mib", type=int, default=1024, help="original bytes of the streaming archive (256 KiB frames)")
    ap.add_argument("--compress-size-mib", type=int, default=1024, help="bytes of mixed data per GPU for the CompressBuffer leg")
    return ap.parse_args()


# ---------------------------------------------------------------- workload
def build_archive(size, frame_size, level, seed):
    """Synthetic text + reference compressor (all host threads via the oracle/_ref harness). Cached in /tmp so
    the reference arm and this arm, run back to back on one box, share the exact same archive."""
    import refzra
    from zra_b200 import synth

    tag = f"/tmp/zra_bench_{size}_{frame_size}_{level}_{seed}"
    if os.path.exists(tag + ".zra") and os.path.exists(tag + ".raw"):
        return np.fromfile(tag + ".raw", dtype=np.uint8), np.fromfile(tag + ".zra", dtype=np.uint8)
    if not refzra.have_ref():
        raise RuntimeError("oracle/_ref/libzra_ref.so is missing: run __graft_entry__.build() where /root/reference exists")
    data = synth.text(size, seed=seed)
    archive = refzra.ref_compress_mt(data, level, frame_size, True)
    try:
        import shutil

        # (only while /tmp keeps plenty of room: 8 ranks x 8 GiB shards must not fill the disk the JSON line goes to)
        if shutil.disk_usage("/tmp").free > 4 * (data.size + archive.size) + (8 << 30) and data.size <= (2 << 30):
            data.tofile(tag + ".raw")
            archive.tofile(tag + ".zra")
    except OSError:
        for ext in (".raw", ".zra"):
            try:
                os.remove(tag + ext)
            except OSError:
                pass
    return data, archive


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        self.index = index
        self.thread = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append((float(out[0]), float(out[1])))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons)}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(self.reasons),
                "samples": len(sm)}


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def kernel_source_hash():
    """sha256 over the decode kernel sources: tools/summarise_profiles.py stamps it into every ncu summary, and a summary
    whose stamp differs from the sources of this tree describes OTHER kernels (it is refused, not silently used)."""
    import glob
    import hashlib

    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "zra_b200", "csrc", "decode_*.cu*")) + glob.glob(os.path.join(ROOT, "zra_b200", "csrc", "*reader.cuh"))
                    + glob.glob(os.path.join(ROOT, "zra_b200", "csrc", "entropy.cuh"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of `kernel`, from the newest committed
    `ncu --set full` summary under profiles/ (tools/summarise_profiles.py; captured on the bench archive with
    ZRA_B200_CHUNKS=1, i.e. one launch = all frames, like the per-kernel timing pass). The kernel column is matched by
    PREFIX (ncu names templates `void k_seq_decode<0>`), and only a summary stamped with the current kernel sources counts.
    Returns (bytes, file, note); bytes is None when no current capture exists."""
    import csv
    import glob

    want = kernel_source_hash()
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_summary.csv")), key=os.path.getmtime)
    stale = None
    for f in reversed(files):
        try:
            lines = open(f).read().splitlines()
            stamp = [l.split("=", 1)[1].strip() for l in lines if l.startswith("# source_sha256=")]
            rows = list(csv.reader([l for l in lines if not l.startswith("#")]))
            head = rows[0]
            cols = [i for i, name in enumerate(head) if name.replace("void ", "").split("<")[0].strip() == kernel]
            if not cols:
                continue
            if not stamp or stamp[0] != want:
                stale = stale or os.path.relpath(f, ROOT)
                continue
            tot = 0.0
            for r in rows[1:]:
                if r and r[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[1]]
                    tot += sum(float(r[c]) for c in cols) * scale
            return int(tot), os.path.relpath(f, ROOT), None
        except Exception:
            continue
    return None, None, (f"newest capture {stale} was taken from other kernel sources" if stale else "no ncu --set full summary under profiles/")


# ---------------------------------------------------------------- CPU reference arm
def cpu_reference(archive, data, threads, repeats=3):
    """The reference's CPU decompression of the SAME archive: best of `repeats`."""
    import ctypes as C

    import refzra

    L = refzra.ref()
    out = np.zeros(data.size, np.uint8)  # pre-faulted
    st = (C.c_int * 2)()
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        if threads == 1:
            s = L.ZraDecompressBuffer(refzra._p(archive), archive.size, refzra._p(out))
            assert s.zra == 0
        else:
            rc = L.ref_decompress_mt(refzra._p(archive), archive.size, refzra._p(out), out.size, threads, st)
            assert rc == 0, list(st)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    assert np.array_equal(out, data)
    return data.size / best / 1e9, best


def workload_text(args, world):
    """config.workload, shared by both arms (the driver compares the strings)."""
    if world == 1:
        return (f"DecompressBuffer, {args.size_mib} MiB Zipf-text archive, {args.frame_size} B frames, level {args.level}, checksums, "
                "written by the reference compressor")
    return (f"one {world * args.size_mib} MiB Zipf-text archive ({args.frame_size} B frames, level {args.level}, checksums, "
            f"reference-compressed), frames sharded contiguously over {world} GPUs ({args.size_mib} MiB per GPU), "
            "each rank decodes its frame range (ZraCudaDecompressFrames); no data-path collective")


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref = the unmodified reference compiled here), all
    host threads, on the GPU arm's workload: at N = 1 the same archive; at N > 1 the same N shards (rank r's archive is
    seed 7 + r), decoded one after another — the host has one set of cores however many GPUs the box has. EXACTLY
    --steps timed steps after --warmup untimed ones; a step is one pass over the whole workload. `value` is the BEST
    step (the statistic of the GPU arm's cpu_baseline leg), the mean is reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    size = args.size_mib << 20
    shards = [build_archive(size, args.frame_size, args.level, seed=7 + r) for r in range(world)]
    threads = os.cpu_count() or 1

    def step():
        t = 0.0
        for data, archive in shards:
            t += cpu_reference(archive, data, threads, repeats=1)[1]
        return t

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    best, mean = min(times), sum(times) / len(times)
    total = world * size
    value = total / best / 1e9
    line = {
        "impl": "reference", "metric": "decompress GB/s", "value": round(value, 4), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(best * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "statistic": "best of the timed steps", "mean": {"value": round(total / mean / 1e9, 4), "ms_per_step": round(mean * 1e3, 3)},
        "config": {"workload": workload_text(args, world), "parallelism": f"frame-shard x{world}",
                   "archive_bytes": int(shards[0][1].size), "original_bytes": size,
                   "frames": (size + args.frame_size - 1) // args.frame_size, "host_threads": threads,
                   "l2": "inputs larger than L2 (archive + output >> 126 MB); no flush needed",
                   "value_definition": "original (decompressed) bytes per second, all GPUs"},
        "cpu_baseline": {"value": round(value, 4), "unit": "GB/s", "cores": threads, "kind": "reference",
                         "sample": f"the whole workload ({world} x {args.size_mib} MiB), {threads} threads each driving its own zra::Decompressor "
                                   "over a disjoint frame range (harness-level parallelism; the reference itself is single-threaded)"},
        "e2e": {"value": round(value, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- random access over the SHARDED archive (configs[3])
def run_ra_sharded(args, torch, dist, ctx, rank, world, d_shard, shard_archive_size, d_data, shard_bytes, fs, rsz):
    """BASELINE configs[3] as it is stated: ONE archive of world x shard_bytes whose frames shard contiguously over the
    GPUs; every rank issues shard_bytes / fs uniform random 4 KiB reads into the WHOLE archive (1 M reads into 16 GiB at
    8 x 2 GiB). Reads are routed to the owners of their bytes with all_to_all_single over NCCL (zra_b200/shard.py
    ShardedReads), served by ZraCudaDecompressRABatch from the owner's resident shard, and the bytes return the same way.
    The owner checks EVERY piece it serves against its original bytes; the requester checks a per-read checksum."""
    from zra_b200 import shard

    total = world * shard_bytes
    count = shard_bytes // fs
    lo = rank * shard_bytes
    rng = np.random.default_rng(4242 + rank)
    offs = torch.from_numpy(rng.integers(0, total - rsz - 1, count).astype(np.int64)).cuda()
    stream = torch.cuda.current_stream()
    checked = [0]
    verify = [True]

    def serve(abs_off, sizes):
        m = int(abs_off.numel())
        local = (abs_off - lo).contiguous()
        sz32 = sizes.to(torch.int32).contiguous()
        starts = (torch.cumsum(sizes, 0) - sizes).contiguous()
        res = torch.empty(int(sizes.sum().item()), dtype=torch.uint8, device="cuda")
        uniform = bool((sizes == rsz).all().item())
        if uniform:
            ctx.decompress_ra_batch(d_shard.data_ptr(), shard_archive_size, local.data_ptr(), m, res.data_ptr(), uniform_size=rsz,
                                    stream=stream.cuda_stream)
        else:
            ctx.decompress_ra_batch(d_shard.data_ptr(), shard_archive_size, local.data_ptr(), m, res.data_ptr(), d_sizes=sz32.data_ptr(),
                                    d_out_offsets=starts.data_ptr(), max_size=rsz, stream=stream.cuda_stream)
        if verify[0] and uniform:
            want = d_data[(local.view(-1, 1) + torch.arange(rsz, device="cuda").view(1, -1)).reshape(-1)]
            assert torch.equal(want, res), "a served read differs from the shard's original bytes"
            checked[0] += m
        return res

    sr = shard.ShardedReads(total, shard_bytes, rsz, device="cuda")
    got = sr.read(offs, serve)
    # requester-side check: the bytes of reads that landed in this rank's own shard are known here
    mine = (offs >= lo) & (offs + rsz <= lo + shard_bytes)
    idx = torch.nonzero(mine).view(-1)
    want = d_data[((offs[idx] - lo).view(-1, 1) + torch.arange(rsz, device="cuda").view(1, -1)).reshape(-1)].view(-1, rsz)
    assert torch.equal(got[idx], want), "a routed read came back different"
    served = torch.tensor([checked[0]], dtype=torch.int64, device="cuda")
    dist.all_reduce(served)
    verify[0] = False
    steps = max(3, min(args.steps, 10))
    sr.read(offs, serve)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        sr.read(offs, serve)
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    return {"metric": "random 4 KiB reads/s, one archive sharded over the GPUs", "value": round(world * count / dt, 1), "unit": "reads/s",
            "ms_per_step": round(dt * 1e3, 3), "reads_per_step": world * count, "archive_original_bytes": int(total),
            "reads_verified_by_owner": int(served.item()),
            "config": {"workload": f"{world * count} uniform random 4 KiB reads into ONE {total >> 20} MiB archive (16384 B frames) sharded over "
                                   f"{world} GPUs ({shard_bytes >> 20} MiB each): all_to_all_single (NCCL) of the offsets to the owning shards, "
                                   "ZraCudaDecompressRABatch there, all_to_all_single of the 4 KiB results back; host-timed, max over ranks"}}


# ---------------------------------------------------------------- batched random access (BASELINE configs[3] shape)
def run_ra(args, torch, dist, ctx, rank, world, peak):
    """4 KiB reads at uniform random offsets into a 16 KiB-frame text archive, one read per frame on average
    (configs[3]: 1M reads over a 16 GiB / 1M-frame archive; here ra-size-mib per GPU at the same density).
    Device-resident: archive, offsets and results in HBM. Returns the dict for the JSON line."""
    fs, rsz = 16384, 4096
    size = (2048 if world == 8 and args.ra_size_mib == 1024 else args.ra_size_mib) << 20   # 8 x 2 GiB = configs[3]'s 16 GiB
    data, archive = build_archive(size, fs, 3, seed=107 + rank)
    count = size // fs
    rng = np.random.default_rng(42 + rank)
    offs = rng.integers(0, size - rsz - 1, count).astype(np.uint64)
    d_in = torch.zeros(archive.size + 64, dtype=torch.uint8, device="cuda")
    d_in[: archive.size] = torch.from_numpy(archive).cuda()
    d_off = torch.from_numpy(offs.view(np.int64)).cuda()
    d_out = torch.empty(count * rsz, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def step():
        return ctx.decompress_ra_batch(d_in.data_ptr(), archive.size, d_off.data_ptr(), count, d_out.data_ptr(), uniform_size=rsz,
                                       stream=stream.cuda_stream)

    unique = step()
    # EVERY read is compared with the original (gathered on the device)
    d_data = torch.from_numpy(data).cuda()
    want = d_data[(d_off.view(-1, 1) + torch.arange(rsz, device="cuda").view(1, -1)).reshape(-1)]
    assert torch.equal(want, d_out), "random-access result differs from the original"
    del want
    step()
    steps = max(3, min(args.steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    # end to end: offsets from pinned host memory, results back to pinned host memory, every step
    h_off = torch.from_numpy(offs.view(np.int64)).pin_memory()
    h_out = torch.empty(count * rsz, dtype=torch.uint8).pin_memory()
    # (one batch per step: splitting it so that a sub-batch's reads go down while the next decodes is SLOWER — 5.1 M
    # against 7.0 M reads/s with four sub-batches — because a batch's decode is latency-bound and takes about as long for
    # 16 384 reads as for 65 536)
    def e2e_step():
        d_off.copy_(h_off, non_blocking=True)
        step()
        h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
    h_out.zero_()
    e2e_step()
    assert np.array_equal(h_out.numpy().reshape(count, rsz), data[(offs.astype(np.int64).reshape(-1, 1) + np.arange(rsz).reshape(1, -1))]), \
        "random-access result (host path) differs from the original"
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    dt = (time.perf_counter() - t0) / steps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    # algorithmic bytes: compressed bytes of the unique frames touched + bytes delivered
    comp = archive.size * (unique / (size // fs))
    alg = comp + count * rsz
    out = {"metric": "random 4 KiB reads/s", "value": round(world * count / (ms / 1e3), 1), "unit": "reads/s", "ms_per_step": round(ms, 4),
           "reads_per_step": world * count, "unique_frames_per_step": world * int(unique),
           "config": {"workload": f"batched DecompressRA, {count} uniform random 4 KiB reads per GPU into a {size >> 20} MiB "
                                  "Zipf-text archive, 16384 B frames, level 3 (configs[3] density: one read per frame)"},
           "e2e": {"value": round(world * count / dt, 1), "unit": "reads/s", "h2d_bytes_per_step": int(offs.nbytes),
                   "d2h_bytes_per_step": int(count * rsz)},
           "roofline": {"bound": "hbm", "achieved": round(alg / (ms / 1e3) / 1e9, 2), "peak": peak, "unit": "GB/s",
                        "frac": round(alg / (ms / 1e3) / 1e9 / peak, 5), "algorithmic_bytes_per_step": int(alg)}}
    if world > 1:
        out["sharded"] = run_ra_sharded(args, torch, dist, ctx, rank, world, d_in, archive.size, d_data, size, fs, rsz)
    del d_data
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            import ctypes as C

            import refzra

            L = refzra.ref()
            threads = os.cpu_count() or 1
            sample = min(count, 32768)
            o = np.ascontiguousarray(offs[:sample])
            res = np.zeros(sample * rsz, np.uint8)
            st = (C.c_int * 2)()
            t0 = time.perf_counter()
            rc = L.ref_ra_mt(refzra._p(archive), archive.size, o.ctypes.data_as(C.POINTER(C.c_uint64)), sample, rsz, refzra._p(res), threads, st)
            dtc = time.perf_counter() - t0
            assert rc == 0 and np.array_equal(res.reshape(sample, rsz)[5], data[int(o[5]): int(o[5]) + rsz])
            out["cpu_baseline"] = {"value": round(sample / dtc, 1), "unit": "reads/s", "cores": threads, "kind": "reference",
                                   "sample": f"the first {sample} reads, {threads} host threads each with its own zra::Decompressor"}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    del d_in, d_out
    return out


# ---------------------------------------------------------------- CompressBuffer (BASELINE configs[2] shape)
def run_compress(args, torch, dist, ctx, rank, world, peak):
    """CompressBuffer of mixed synthetic data (64 KiB runs of Zipf text alternating with incompressible bytes), 64 KiB
    frames, levels 3 and 1, checksums. configs[2] is 8 GiB sharded over the GPUs of the box; here compress-size-mib per
    GPU (1 GiB = the per-GPU shard of the 8-GPU case). The archive is verified by the REFERENCE decoder (oracle/_ref)."""
    import ctypes as C

    import refzra
    import zra_b200
    from zra_b200 import synth

    fs = 65536
    size = args.compress_size_mib << 20
    data = synth.mixed(size, period=fs, seed=7 + rank)
    d_in = torch.from_numpy(data).cuda()
    cap = zra_b200_cap(size, fs)
    d_out = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    out = {"metric": "compress GB/s", "unit": "GB/s", "config": {
        "workload": f"CompressBuffer, {args.compress_size_mib} MiB per GPU of mixed data (64 KiB text / 64 KiB incompressible "
                    "alternating), 65536 B frames, checksums", "value_definition": "input (uncompressed) bytes per second, all GPUs"},
        "levels": {}}
    if world > 1:
        out["config"]["workload"] = (f"CompressBuffer of ONE {world * args.compress_size_mib} MiB mixed input (64 KiB text / 64 KiB incompressible "
                                     f"alternating), 65536 B frames, checksums, frames sharded contiguously over {world} GPUs: every rank compresses its "
                                     "frame range (ZraCudaCompressFrames), the per-frame sizes are all-gathered over NCCL (= the scan of the per-shard "
                                     "totals), every rank stitches the header (ZraShardBuildHeader); the stitched archive is decoded by the reference")
    for level in (3, 1):
        if world == 1:
            def step():
                return ctx.compress_buffer(d_in.data_ptr(), size, d_out.data_ptr(), cap, level=level, frame_size=fs, checksum=True,
                                           stream=stream.cuda_stream)
        else:
            from zra_b200 import shard

            frames_total = world * (size // fs)
            sharded = {}

            def step():
                # the sharded path's whole step: compress the shard, exchange the frame sizes (the one collective), build the header
                sizes, produced = shard.compress_shard(ctx, d_in.data_ptr(), size, fs, level, True, d_out.data_ptr(), cap, stream.cuda_stream)
                all_sizes, base, total = shard.exchange_frame_sizes(sizes, frames_total)
                header = shard.build_header(world * size, fs, all_sizes)
                sharded.update(produced=produced, base=base, total=total, header=header)
                return header.size + total    # bytes of the whole stitched archive
        t0 = time.perf_counter()
        n = step()
        torch.cuda.synchronize()
        first = time.perf_counter() - t0
        if world == 1:
            archive = d_out[:n].cpu().numpy()
            if rank == 0 and refzra.have_ref():  # the unmodified reference must decode what the GPU wrote
                back = np.zeros(size, np.uint8)
                st = (C.c_int * 2)()
                rc = refzra.ref().ref_decompress_mt(refzra._p(archive), archive.size, refzra._p(back), back.size, os.cpu_count() or 1, st)
                assert rc == 0 and np.array_equal(back, data), f"reference decoder rejects the GPU archive (level {level}): {list(st)}"
        elif level == 3:
            # gather the shards' frames on rank 0 (NCCL), stitch header + frames, and let the REFERENCE decode the whole archive;
            # every rank's original bytes are compared through a checksum of checksums (sum of bytes per 1 MiB block)
            produced = sharded["produced"]
            lens = torch.zeros(world, dtype=torch.int64, device="cuda")
            lens[rank] = produced
            dist.all_reduce(lens)
            lens = [int(x) for x in lens.tolist()]
            block_sums = torch.from_numpy(data).cuda().view(-1, 1 << 20).sum(dim=1, dtype=torch.int64)
            all_sums = [torch.empty_like(block_sums) for _ in range(world)] if rank == 0 else None
            dist.gather(block_sums, all_sums, dst=0)
            if rank == 0:
                whole = torch.empty(sharded["header"].size + sum(lens), dtype=torch.uint8, device="cuda")
                whole[: sharded["header"].size] = torch.from_numpy(sharded["header"]).cuda()
                at = sharded["header"].size
                whole[at: at + lens[0]] = d_out[: lens[0]]
                at += lens[0]
                for r in range(1, world):
                    dist.recv(whole[at: at + lens[r]], src=r)
                    at += lens[r]
                stitched = whole.cpu().numpy()
                del whole
                if refzra.have_ref():
                    back = np.zeros(world * size, np.uint8)
                    st = (C.c_int * 2)()
                    rc = refzra.ref().ref_decompress_mt(refzra._p(stitched), stitched.size, refzra._p(back), back.size, os.cpu_count() or 1, st)
                    assert rc == 0, f"reference decoder rejects the stitched sharded archive: {list(st)}"
                    assert np.array_equal(back[:size], data), "shard 0 of the stitched archive differs from its input"
                    got = torch.from_numpy(back).view(world, -1, 1 << 20).sum(dim=2, dtype=torch.int64)
                    for r in range(world):
                        assert torch.equal(got[r], all_sums[r].cpu()), f"shard {r} of the stitched archive differs from its input"
                    out["sharded_archive_verified"] = {"by": "reference decoder (oracle/_ref), all host threads", "bytes": int(stitched.size),
                                                       "original_bytes": int(world * size)}
                    del back
            else:
                dist.send(d_out[: lens[rank]], dst=0)
            dist.barrier()
        steps = max(1, min(args.steps, int(3.0 / max(first, 1e-3))))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / steps
        alg = size + (n if world == 1 else sharded["produced"])   # per GPU
        res = {"value": round(world * size / (ms / 1e3) / 1e9, 3), "ms_per_step": round(ms, 3), "steps": steps, "archive_bytes": int(n),
               "ratio": round((world * size if world > 1 else size) / n, 4),
               "roofline": {"bound": "hbm", "achieved": round(alg / (ms / 1e3) / 1e9, 2), "peak": peak, "unit": "GB/s",
                            "frac": round(alg / (ms / 1e3) / 1e9 / peak, 6), "algorithmic_bytes_per_step": int(alg)}}
        if rank == 0 and world == 1:
            # end to end: ZraCompressBuffer with pinned host buffers
            L = zra_b200.lib()
            h_in = torch.from_numpy(data).pin_memory()
            h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
            osz = C.c_size_t(0)
            def e2e_step():
                st = L.ZraCompressBuffer(C.c_void_p(h_in.data_ptr()), size, C.c_void_p(h_out.data_ptr()), C.byref(osz), level, fs, True, None, 0)
                assert st.zra == 0, (st.zra, st.zstd)
            e2e_step()
            k = max(1, min(steps, 3))
            t0 = time.perf_counter()
            for _ in range(k):
                e2e_step()
            dt = (time.perf_counter() - t0) / k
            res["e2e"] = {"value": round(size / dt / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": size, "d2h_bytes_per_step": int(osz.value),
                          "api": "ZraCompressBuffer (host pointers, pinned)"}
            del h_in, h_out
            if not args.no_cpu_baseline and refzra.have_ref():
                # the reference compressor on a bounded sample (first 256 MiB) with all host threads: ratio and GB/s
                sample = min(size, 256 << 20)
                threads = os.cpu_count() or 1
                t0 = time.perf_counter()
                ref_arch = refzra.ref_compress_mt(data[:sample], level, fs, True, threads)
                dtc = time.perf_counter() - t0
                gpu_sample = ctx.compress_buffer(d_in.data_ptr(), sample, d_out.data_ptr(), cap, level=level, frame_size=fs, checksum=True,
                                                 stream=stream.cuda_stream)
                res["ratio_vs_reference"] = {"sample_bytes": sample, "reference_archive_bytes": int(ref_arch.size),
                                             "gpu_archive_bytes": int(gpu_sample),
                                             "size_delta": round(gpu_sample / ref_arch.size - 1.0, 5)}
                res["cpu_baseline"] = {"value": round(sample / dtc / 1e9, 4), "unit": "GB/s", "cores": threads, "kind": "reference",
                                       "sample": f"the first {sample >> 20} MiB, {threads} host threads each with its own zra::Compressor "
                                                 "(includes stitching the seek table)"}
        out["levels"][f"L{level}"] = res
    if world == 1:
        # frames above 64 KiB take the frame-cooperative matcher with 32-bit tables (k_enc_match_cta_big): its own number
        try:
            fs2 = 262144
            cap2 = zra_b200_cap(size, fs2)
            d_out2 = torch.empty(cap2 + 64, dtype=torch.uint8, device="cuda")
            def step2():
                return ctx.compress_buffer(d_in.data_ptr(), size, d_out2.data_ptr(), cap2, level=3, frame_size=fs2, checksum=True,
                                           stream=stream.cuda_stream)
            n2 = step2()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step2()
            e1.record(stream)
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1)
            out["frames_256KiB_L3"] = {"value": round(size / (ms2 / 1e3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(ms2, 3),
                                       "ratio": round(size / n2, 4), "note": "262144 B frames: k_enc_match_cta_big (one CTA per frame, 32-bit tables in shared memory)"}
            del d_out2
        except Exception as e:  # noqa: BLE001
            out["frames_256KiB_L3"] = {"value": None, "error": repr(e)}
    out["value"] = out["levels"]["L3"]["value"]
    del d_in, d_out
    return out


# ---------------------------------------------------------------- FullDecompressor streaming (BASELINE configs[4] shape)
def run_streaming(args, torch, dist, rank, world):
    """zra::FullDecompressor over the C ABI (ZraCreateFullDecompressor / ZraDecompressWithFullDecompressor): an archive of
    256 KiB frames is streamed through the reference's read-callback model (the callback memcpy's from a host copy of
    the archive, once per Decompress call, synchronously on the caller's thread) into a pinned 256 MiB output buffer,
    call after call until it returns 0. Host pointers in, host pointers out: uploads and downloads are inside the timing."""
    import ctypes as C

    import zra_b200
    from zra_b200 import binding

    fs = 262144
    size = args.stream_size_mib << 20
    data, archive = build_archive(size, fs, 3, seed=207 + rank)
    base = archive.ctypes.data
    calls = [0]

    def cb(offset, nbytes, buf):
        calls[0] += 1
        C.memmove(buf, base + offset, nbytes)

    reader = binding.READ_FN(cb)
    L = zra_b200.lib()
    out = torch.empty(min(size, 256 << 20), dtype=torch.uint8).pin_memory()
    result = np.empty(size, np.uint8)

    def one_pass(keep):
        h = C.c_void_p()
        st = L.ZraCreateFullDecompressor(C.byref(h), reader, 0)
        assert st.zra == 0, (st.zra, st.zstd)
        pos = 0
        n = C.c_size_t(0)
        while True:
            st = L.ZraDecompressWithFullDecompressor(h, C.c_void_p(out.data_ptr()), out.numel(), C.byref(n))
            assert st.zra == 0, (st.zra, st.zstd)
            if not n.value:
                break
            if keep:
                result[pos: pos + n.value] = out.numpy()[: n.value]
            pos += n.value
        L.ZraDeleteFullDecompressor(h)
        return pos

    assert one_pass(True) == size and np.array_equal(result, data), "streamed output differs from the original"
    # the same archive, device-resident, through ZraCudaDecompressBuffer: the 256 KiB-frame decode number (configs[4]'s shape)
    resident = None
    try:
        ctxd = zra_b200.CudaContext(torch.cuda.current_device())
        d_a = torch.zeros(archive.size + 64, dtype=torch.uint8, device="cuda")
        d_a[: archive.size] = torch.from_numpy(archive).cuda()
        d_o = torch.empty(size, dtype=torch.uint8, device="cuda")
        stc = torch.cuda.current_stream()
        for _ in range(3):
            ctxd.decompress_buffer(d_a.data_ptr(), archive.size, d_o.data_ptr(), size, stc.cuda_stream)
        assert torch.equal(d_o, torch.from_numpy(data).cuda())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = max(3, min(args.steps, 10))
        e0.record(stc)
        for _ in range(k):
            ctxd.decompress_buffer(d_a.data_ptr(), archive.size, d_o.data_ptr(), size, stc.cuda_stream)
        e1.record(stc)
        torch.cuda.synchronize()
        msd = e0.elapsed_time(e1) / k
        resident = {"value": round(size / (msd / 1e3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(msd, 4),
                    "note": "ZraCudaDecompressBuffer of the same 262144 B-frame archive, resident in HBM (per GPU)"}
        del d_a, d_o, ctxd
    except Exception as e:  # noqa: BLE001
        resident = {"value": None, "error": repr(e)}
    steps = max(2, min(args.steps, 5))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass(False)
    dt = (time.perf_counter() - t0) / steps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    res = {"metric": "FullDecompressor streaming GB/s", "value": round(world * size / dt / 1e9, 3), "unit": "GB/s",
           "ms_per_pass": round(dt * 1e3, 3),
           "config": {"workload": f"zra::FullDecompressor, {args.stream_size_mib} MiB Zipf-text archive per GPU, 262144 B frames, level 3, "
                                  f"read callback = memcpy from a host copy, {out.numel() >> 20} MiB pinned output buffer per call "
                                  "(configs[4] shape; frame ranges = archives per rank at N > 1)",
                      "read_callbacks_per_pass": calls[0] // (steps + 1)},
           "h2d_bytes_per_pass": int(archive.size), "d2h_bytes_per_pass": int(size), "device_resident_decode_256KiB_frames": resident}
    res["compressor"] = stream_compressor(args, torch, L, data, size)
    if rank == 0 and not args.no_cpu_baseline:
        try:
            import refzra

            R = refzra.ref()
            R.ZraCreateFullDecompressor.argtypes = [C.POINTER(C.c_void_p), binding.READ_FN, C.c_size_t]
            R.ZraCreateFullDecompressor.restype = R.St
            R.ZraDecompressWithFullDecompressor.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
            R.ZraDecompressWithFullDecompressor.restype = R.St
            R.ZraDeleteFullDecompressor.argtypes = [C.c_void_p]
            h = C.c_void_p()
            assert R.ZraCreateFullDecompressor(C.byref(h), reader, 0).zra == 0
            n = C.c_size_t(0)
            host = np.empty(out.numel(), np.uint8)
            host[:] = 0
            t0 = time.perf_counter()
            assert R.ZraDecompressWithFullDecompressor(h, host.ctypes.data, host.size, C.byref(n)).zra == 0
            dtc = time.perf_counter() - t0
            R.ZraDeleteFullDecompressor(h)
            assert np.array_equal(host[: n.value], data[: n.value])
            res["cpu_baseline"] = {"value": round(n.value / dtc / 1e9, 4), "unit": "GB/s", "cores": 1, "kind": "reference",
                                   "sample": f"the first Decompress call ({n.value >> 20} MiB) of the reference's zra::FullDecompressor "
                                             "(single-threaded by construction)"}
        except Exception as e:  # noqa: BLE001
            res["cpu_baseline"] = {"value": None, "sample": f"failed: {e}"}
    return res


def stream_compressor(args, torch, L, data, size):
    """zra::Compressor over the C ABI (ZraCreateCompressor / ZraCompressWithCompressor / ZraGetHeaderWithCompressor): the
    input is fed in 64 MiB pieces from pinned host memory, each call returns that piece's frames in a pinned host buffer
    (the reference's streaming writer model, zra.cpp:304-365); the header comes last. Verified by decoding header +
    frames with ZraDecompressBuffer."""
    import ctypes as C

    fs = 65536
    piece = 64 << 20
    try:
        h_in = torch.from_numpy(data).pin_memory()
        hc = C.c_void_p()

        def one_pass(keep):
            st = L.ZraCreateCompressor(C.byref(hc), size, 3, fs, True, None, 0)
            assert st.zra == 0, (st.zra, st.zstd)
            cap = L.ZraGetOutputBufferSizeWithCompressor(hc, piece)
            h_out = one_pass.out if one_pass.out is not None else torch.empty(cap, dtype=torch.uint8).pin_memory()
            one_pass.out = h_out
            n = C.c_size_t(0)
            parts, total = [], 0
            for off in range(0, size, piece):
                m = min(piece, size - off)
                st = L.ZraCompressWithCompressor(hc, C.c_void_p(h_in.data_ptr() + off), m, C.c_void_p(h_out.data_ptr()), C.byref(n))
                assert st.zra == 0, (st.zra, st.zstd)
                total += n.value
                if keep:
                    parts.append(h_out.numpy()[: n.value].copy())
            hs = L.ZraGetHeaderSizeWithCompressor(hc)
            head = np.empty(hs, np.uint8)
            st = L.ZraGetHeaderWithCompressor(hc, C.c_void_p(head.ctypes.data))
            assert st.zra == 0, (st.zra, st.zstd)
            L.ZraDeleteCompressor(hc)
            return (np.concatenate([head] + parts) if keep else None), hs + total

        one_pass.out = None
        z, zsize = one_pass(True)
        back = np.empty(size, np.uint8)
        st = L.ZraDecompressBuffer(C.c_void_p(z.ctypes.data), z.size, C.c_void_p(back.ctypes.data))
        assert st.zra == 0 and np.array_equal(back, data), "streamed archive does not decode to the input"
        steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(steps):
            one_pass(False)
        dt = (time.perf_counter() - t0) / steps
        return {"metric": "Compressor streaming GB/s", "value": round(size / dt / 1e9, 3), "unit": "GB/s", "ms_per_pass": round(dt * 1e3, 3),
                "ratio": round(size / zsize, 4), "h2d_bytes_per_pass": int(size), "d2h_bytes_per_pass": int(zsize),
                "note": f"zra::Compressor, {size >> 20} MiB per GPU in {piece >> 20} MiB calls, {fs} B frames, level 3, pinned host buffers (per GPU)"}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "error": repr(e)}


def zra_b200_cap(size, frame_size):
    import zra_b200
    return int(zra_b200.GetOutputBufferSize(size, frame_size))


# ---------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import zra_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (zra-b200 has no CPU path)"
    torch.cuda.set_device(local)
    # rank 0's stdout carries exactly ONE JSON line: whatever native libraries print on fd 1 in the meantime
    # (NCCL's version banner, for one) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    size = args.size_mib << 20
    data, archive = build_archive(size, args.frame_size, args.level, seed=7 + rank)
    ctx = zra_b200.CudaContext(local)
    stream = torch.cuda.current_stream()
    d_out = torch.empty(size, dtype=torch.uint8, device="cuda")
    d_ref = torch.from_numpy(data).cuda()
    if world == 1:
        d_in = torch.zeros(archive.size + 64, dtype=torch.uint8, device="cuda")
        d_in[: archive.size] = torch.from_numpy(archive).cuda()
        sharded_bytes = archive.size

        def step():
            ctx.decompress_buffer(d_in.data_ptr(), archive.size, d_out.data_ptr(), size, stream.cuda_stream)
    else:
        # ONE archive of world x size bytes whose frames shard contiguously across the ranks: rank r made
        # frames [r*F/W, (r+1)*F/W); the per-frame sizes are all-gathered over NCCL (the scan of the per-shard totals
        # gives each shard its base offset), every rank stitches the same header and holds header + its own frames.
        from common import parse_header, seek_table
        from zra_b200 import shard

        h = parse_header(archive)
        sizes = np.diff(seek_table(archive))
        frames_total = world * (size // args.frame_size)
        assert size % args.frame_size == 0
        all_sizes, base, total = shard.exchange_frame_sizes(sizes, frames_total)
        header = shard.build_header(world * size, args.frame_size, all_sizes)
        sharded_bytes = header.size + total
        d_in = torch.zeros(sharded_bytes + 64, dtype=torch.uint8, device="cuda")
        d_in[: header.size] = torch.from_numpy(header).cuda()
        payload = archive[h["size"]:]
        d_in[header.size + base: header.size + base + payload.size] = torch.from_numpy(payload).cuda()
        f0, f1 = shard.frame_range(frames_total, rank, world)

        def step():
            ctx.decompress_frames(d_in.data_ptr(), sharded_bytes, f0, f1 - f0, d_out.data_ptr(), size, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    assert torch.equal(d_out, d_ref), "decompressed bytes differ from the original"

    # ---- timed region: device-resident
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        launches = ctx.launch_count() - launches0
        # the timed region lasts tens of milliseconds, one nvidia-smi poll at best: keep the same load running (untimed)
        # for about a second more so that the clock / throttle samples describe this workload, not an idle GPU
        t_hold = time.perf_counter()
        while time.perf_counter() - t_hold < 1.0:
            step()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * size * args.steps / (ms_max / 1e3) / 1e9

    # ---- N > 1: the final gather of configs[4] (every rank ends with the whole decoded archive), NCCL all-gather over
    # NVLink / NVSwitch, timed separately from the decode (bounded by NVLink, reported on its own)
    gather = None
    if world > 1:
        out_all = torch.empty(world * size, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            dist.all_gather_into_tensor(out_all, d_out)
        sums = torch.zeros(world, dtype=torch.int64, device="cuda")
        # (a checksum of 64-bit words: no widened temporary, which at 8 GiB per shard would not fit)
        sums[rank] = d_ref.view(torch.int64).sum()
        dist.all_reduce(sums)
        got = out_all.view(torch.int64).view(world, size // 8).sum(dim=1)
        assert torch.equal(got, sums), "gathered archive differs from the shards' originals"
        assert torch.equal(out_all[rank * size:(rank + 1) * size], d_ref)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gsteps = max(3, min(args.steps, 10))
        barrier()
        g0.record(stream)
        for _ in range(gsteps):
            dist.all_gather_into_tensor(out_all, d_out)
        g1.record(stream)
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1) / gsteps], dtype=torch.float64, device="cuda")
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gms = float(tg.item())
        gather = {"op": "ncclAllGather of the decoded shards (every GPU ends with the whole archive)", "ms": round(gms, 4),
                  "bytes_in_per_gpu": int((world - 1) * size), "GBps_in_per_gpu": round((world - 1) * size / gms / 1e6, 2),
                  "nvlink_peak_GBps_per_direction": 900.0,
                  "decode_then_gather_GBps": round(world * size / ((ms_max / args.steps + gms) / 1e3) / 1e9, 3)}
        # ---- the same result with the gather OVERLAPPED with the decode: the rank's frames are decoded in two slices,
        # straight into their place in the gathered buffer; as soon as a slice is done its bytes go to every peer
        # (batched NCCL send / recv over NVSwitch, into place) while the next slice decodes. Bound: the gather alone.
        try:
            Q = 2   # each slice must still fill the sequence stage (a decode of fewer frames takes as long: the stages are latency-bound)
            fpr = size // args.frame_size
            if fpr % Q == 0:
                qn, qb = fpr // Q, size // Q
                mine = out_all[rank * size:(rank + 1) * size]

                def overlapped():
                    works = []
                    for q in range(Q):
                        ctx.decompress_frames(d_in.data_ptr(), sharded_bytes, f0 + q * qn, qn, mine.data_ptr() + q * qb, qb, stream.cuda_stream)
                        ops = []
                        for peer in range(world):
                            if peer == rank:
                                continue
                            ops.append(dist.P2POp(dist.isend, mine[q * qb:(q + 1) * qb], peer))
                            ops.append(dist.P2POp(dist.irecv, out_all[peer * size + q * qb: peer * size + (q + 1) * qb], peer))
                        works += dist.batch_isend_irecv(ops)
                    for w in works:
                        w.wait()

                out_all.zero_()
                overlapped()
                torch.cuda.synchronize()
                got = out_all.view(torch.int64).view(world, size // 8).sum(dim=1)
                assert torch.equal(got, sums), "overlapped gather: the gathered archive differs from the shards' originals"
                assert torch.equal(mine, d_ref)
                barrier()
                t0 = time.perf_counter()
                for _ in range(gsteps):
                    overlapped()
                torch.cuda.synchronize()
                to = torch.tensor([(time.perf_counter() - t0) / gsteps * 1e3], dtype=torch.float64, device="cuda")
                dist.all_reduce(to, op=dist.ReduceOp.MAX)
                oms = float(to.item())
                gather["overlapped"] = {"what": f"decode in {Q} slices into place, each slice sent to all peers (batched ncclSend/ncclRecv) while the next decodes",
                                        "ms": round(oms, 4), "decode_and_gather_GBps": round(world * size / (oms / 1e3) / 1e9, 3),
                                        "gather_only_bound_GBps": round(world * size / (gms / 1e3) / 1e9, 3),
                                        "fraction_of_gather_only_bound": round(gms / oms, 4)}
        except Exception as e:  # noqa: BLE001
            gather["overlapped"] = {"error": repr(e)}
        del out_all
        torch.cuda.empty_cache()

    # ---- per-kernel breakdown (separate pass with an event after every launch)
    ctx.set_profiling(True)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    prof = ctx.kernel_profile()
    ctx.set_profiling(False)
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items() if v[1]}
    total_kernel_ms = sum(k["ms_per_step"] for k in kernels.values())
    top = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    peak, peak_kind = measured_peak()
    alg_bytes = archive.size + size
    top_ms = kernels[top]["ms_per_step"]
    traffic, traffic_src, traffic_why = ncu_traffic(top)
    top_launches = max(1.0, kernels[top]["launches_per_step"])
    roofline = {
        "bound": "hbm", "kernel": top, "achieved": round(alg_bytes / (top_ms / 1e3) / 1e9, 2), "peak": peak, "peak_kind": peak_kind,
        "unit": "GB/s", "frac": round(alg_bytes / (top_ms / 1e3) / 1e9 / peak, 5), "traffic": traffic,
        "traffic_note": (f"DRAM read+write bytes of ONE launch of {top} from {traffic_src} (ncu --set full on the same archive with "
                         "ZRA_B200_CHUNKS=1: one launch = all frames, the geometry of this timing pass)") if traffic else traffic_why,
        "launches_per_step": top_launches,
        "algorithmic_bytes_per_step": int(alg_bytes), "kernel_ms_per_step": round(top_ms, 4),
        "kernel_share_of_step": round(top_ms / total_kernel_ms, 4),
        # THE number to quote: every algorithmic byte of the step over the whole step's time (the per-kernel `frac`
        # above credits one kernel with all of the step's bytes, as the measurement recipe defines it)
        "whole_step_frac": round(alg_bytes / (ms_max / args.steps / 1e3) / 1e9 / peak, 5),
        "whole_step": {"achieved": round(alg_bytes / (ms_max / args.steps / 1e3) / 1e9, 2),
                       "frac": round(alg_bytes / (ms_max / args.steps / 1e3) / 1e9 / peak, 5)},
        "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in kernels.items()},
    }

    # ---- end to end through the reference-facing C-ABI with pinned host buffers
    h_in = torch.from_numpy(archive).pin_memory()
    h_out = torch.empty(size, dtype=torch.uint8).pin_memory()
    L = zra_b200.lib()
    import ctypes as C

    def e2e_step():
        st = L.ZraDecompressBuffer(C.c_void_p(h_in.data_ptr()), archive.size, C.c_void_p(h_out.data_ptr()))
        assert st.zra == 0, (st.zra, st.zstd)

    for _ in range(2):
        e2e_step()
    assert np.array_equal(h_out.numpy(), data), "e2e output differs from the original"
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * size * e2e_steps / float(t.item()) / 1e9

    ra = None
    if not args.no_ra:
        del d_ref
        torch.cuda.empty_cache()
        ra = run_ra(args, torch, dist, ctx, rank, world, peak)

    comp = None
    if not args.no_compress:
        torch.cuda.empty_cache()
        try:
            comp = run_compress(args, torch, dist, ctx, rank, world, peak)
        except Exception as e:  # never take the headline down
            comp = {"metric": "compress GB/s", "value": None, "unit": "GB/s", "error": repr(e)}

    streaming = None
    if not args.no_streaming:
        try:
            streaming = run_streaming(args, torch, dist, rank, world)
        except Exception as e:  # never take the headline down
            streaming = {"metric": "FullDecompressor streaming GB/s", "value": None, "unit": "GB/s", "error": repr(e)}

    if rank == 0:
        line = {
            "metric": "decompress GB/s", "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_text(args, world),
                       "parallelism": f"frame-shard x{world}",
                       "archive_bytes": int(archive.size), "original_bytes": size, "frames": (size + args.frame_size - 1) // args.frame_size,
                       "l2": "inputs larger than L2 (archive + output >> 126 MB); no flush needed",
                       "value_definition": "original (decompressed) bytes per second, all GPUs"},
            "clocks": dict(clocks.summary(), window="the timed region plus 1 s of the same steps, untimed"),
            "e2e": {"value": round(e2e_value, 3), "unit": "GB/s", "h2d_bytes_per_step": int(archive.size), "d2h_bytes_per_step": size,
                    "api": "ZraDecompressBuffer (host pointers, pinned)", "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if gather is not None:
            line["gather"] = gather
        if ra is not None:
            line["random_access"] = ra
        if comp is not None:
            line["compress"] = comp
        if streaming is not None:
            line["streaming"] = streaming
        if world == 1 and not args.no_cpu_baseline:
            try:
                threads = os.cpu_count() or 1
                v1, t1 = cpu_reference(archive, data, 1, repeats=1)
                vN, tN = cpu_reference(archive, data, threads, repeats=2)
                line["cpu_baseline"] = {"value": round(vN, 4), "unit": "GB/s", "cores": threads, "kind": "reference",
                                        "sample": f"the same {args.size_mib} MiB archive, all {threads} host threads (one zra::Decompressor each)",
                                        "single_thread": {"value": round(v1, 4), "unit": "GB/s", "cores": 1,
                                                          "sample": "zra::DecompressBuffer, faithful single-thread reference"}}
            except Exception as e:  # the baseline must never take the GPU numbers down with it
                line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(json_fd, 1)
    os.close(json_fd)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
