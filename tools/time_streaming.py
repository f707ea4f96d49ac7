"""Per-call timing of zra::FullDecompressor over the C ABI (the bench's streaming leg, dissected): time inside the read
callback and time of every Decompress call. usage: time_streaming.py [size_mib] [frame_size] [out_mib]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import zra_b200  # noqa: E402
from zra_b200 import binding  # noqa: E402

size = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
fs = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
out_mib = int(sys.argv[3]) if len(sys.argv) > 3 else 256
data, archive = bench.build_archive(size, fs, 3, seed=207)
base = archive.ctypes.data
cb_t = []


def cb(offset, nbytes, buf):
    t = time.perf_counter()
    C.memmove(buf, base + offset, nbytes)
    cb_t.append((time.perf_counter() - t, nbytes))


reader = binding.READ_FN(cb)
L = zra_b200.lib()
out = torch.empty(min(size, out_mib << 20), dtype=torch.uint8).pin_memory()
for trial in range(3):
    h = C.c_void_p()
    assert L.ZraCreateFullDecompressor(C.byref(h), reader, 0).zra == 0
    cb_t.clear()
    calls = []
    n = C.c_size_t(0)
    t0 = time.perf_counter()
    while True:
        t = time.perf_counter()
        st = L.ZraDecompressWithFullDecompressor(h, C.c_void_p(out.data_ptr()), out.numel(), C.byref(n))
        assert st.zra == 0
        calls.append(round((time.perf_counter() - t) * 1e3, 2))
        if not n.value:
            break
    total = time.perf_counter() - t0
    L.ZraDeleteFullDecompressor(h)
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("ZRA_B200")}, "GBps": round(size / total / 1e9, 2), "call_ms": calls,
                  "callback_ms": [round(t * 1e3, 2) for t, nb in cb_t if nb], "callback_GBps": round(sum(nb for t, nb in cb_t) / max(1e-9, sum(t for t, nb in cb_t)) / 1e9, 2)}))
