"""Pinned-memory PCIe bandwidth with ALL ranks of a torchrun launch copying at the same time — the box's aggregate
host<->device ceiling, which bounds the host-pointer API (e2e) at N GPUs. Reports, per direction and for both at once,
every rank's GB/s and the aggregate; also the NUMA node of each GPU, the CPU affinity of each rank and the memory
policy, because a pinned buffer on the far socket halves the link.
usage: torchrun --nproc-per-node N tools/pcie_probe_ranks.py [mib]  -> one JSON line on rank 0"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = mib << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
h_a.fill_(1); h_b.fill_(2)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(mode, reps=4):
    """All ranks start together (barrier), copy `reps` times; returns this rank's GB/s (per direction)."""
    def once():
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_a.copy_(h_a, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)
    once(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return n * reps / dt / 1e9


res = {}
for mode in ("h2d", "d2h", "both"):
    v = torch.tensor([run(mode)], dtype=torch.float64, device="cuda")
    if world > 1:
        allv = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
        vals = [round(float(x.item()), 2) for x in allv]
    else:
        vals = [round(float(v.item()), 2)]
    res[mode] = {"per_rank_GBps": vals, "aggregate_GBps": round(sum(vals), 2)}


def numa_of_gpu(i):
    try:
        import subprocess
        bus = subprocess.run(["nvidia-smi", "-i", str(i), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        p = "/sys/bus/pci/devices/" + bus.lower()[4:] + "/numa_node"
        return int(open(p).read())
    except Exception:
        return None


info = {"rank": rank, "gpu_numa": numa_of_gpu(local), "affinity_cpus": len(os.sched_getaffinity(0))}
if world > 1:
    infos = [None] * world
    dist.all_gather_object(infos, info)
else:
    infos = [info]
if rank == 0:
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    except Exception:
        nodes = []
    print(json.dumps({"ranks": world, "mib_per_copy": mib, "host_cpus": os.cpu_count(), "numa_nodes": nodes, "per_rank": infos, **res}), flush=True)
if world > 1:
    dist.destroy_process_group()
