"""Times device-resident CompressBuffer calls on the bench's mixed data (tuning; run under ncu for a launch list).
usage: time_compress.py [size_mib] [frame_size] [level] [steps] [kind=mixed|text] -> one JSON line"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import zra_b200  # noqa: E402
from zra_b200 import synth  # noqa: E402

size_mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fs = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
level = int(sys.argv[3]) if len(sys.argv) > 3 else 3
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
kind = sys.argv[5] if len(sys.argv) > 5 else "mixed"
size = size_mib << 20
data = synth.mixed(size, period=65536, seed=7) if kind == "mixed" else synth.text(size, seed=7)
ctx = zra_b200.CudaContext(0)
d_in = torch.from_numpy(data).cuda()
cap = zra_b200.GetOutputBufferSize(size, fs)
d_out = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()


def step():
    return ctx.compress_buffer(d_in.data_ptr(), size, d_out.data_ptr(), cap, level=level, frame_size=fs, checksum=True,
                               stream=st.cuda_stream)


n = step()
torch.cuda.synchronize()
archive = d_out[:n].cpu().numpy()
ok = None
try:
    import refzra
    if refzra.have_ref():
        ok = bool(np.array_equal(refzra.ref_decompress(archive), data))
except Exception as e:  # noqa: BLE001
    ok = repr(e)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(steps):
    step()
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
env = {k: v for k, v in os.environ.items() if k.startswith("ZRA_B200_")}
print(json.dumps({"kind": kind, "size_mib": size_mib, "frame": fs, "level": level, "ok": ok, "archive_bytes": int(n),
                  "ratio": round(size / n, 4), "ms_per_step": round(ms, 3), "GBps": round(size / ms / 1e6, 3), "env": env}), flush=True)
