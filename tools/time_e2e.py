"""Times the host-pointer C-ABI (ZraDecompressBuffer, pinned buffers) on the bench archive under the current environment.
usage: time_e2e.py [size_mib] [frame_size] [steps] [tag]  -> one JSON line"""
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

size_mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
frame = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
tag = sys.argv[4] if len(sys.argv) > 4 else ""
data, archive = bench.build_archive(size_mib << 20, frame, 3, seed=7)
h_in = torch.from_numpy(archive).pin_memory()
h_out = torch.empty(data.size, dtype=torch.uint8).pin_memory()
L = zra_b200.lib()


def step():
    st = L.ZraDecompressBuffer(C.c_void_p(h_in.data_ptr()), archive.size, C.c_void_p(h_out.data_ptr()))
    assert st.zra == 0, (st.zra, st.zstd)


for _ in range(2):
    step()
ok = bool(np.array_equal(h_out.numpy(), data))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    step()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / steps * 1e3
env = {k: v for k, v in os.environ.items() if k.startswith("ZRA_B200_")}
print(json.dumps({"tag": tag, "ok": ok, "ms_per_step": round(ms, 3), "GBps": round(data.size / ms / 1e6, 2), "env": env}), flush=True)
