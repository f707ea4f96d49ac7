"""Pinned-memory PCIe bandwidth of the box (what bounds the host-pointer API): H2D, D2H, and both at once.
usage: pcie_probe.py [mib]  -> one JSON line"""
import json
import sys

import torch

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        s1.synchronize(); s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def h2d():
    s1.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)


def d2h():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)


for f in (h2d, d2h, both):
    f()
out = {"mib": mib, "h2d_GBps": round(n / timed(h2d) / 1e6, 2), "d2h_GBps": round(n / timed(d2h) / 1e6, 2)}
t = timed(both)
out["duplex_each_GBps"] = round(n / t / 1e6, 2)
print(json.dumps(out), flush=True)
