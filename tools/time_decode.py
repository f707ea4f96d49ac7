"""Times device-resident decodes of the bench archive under the current environment (tuning sweeps).
usage: time_decode.py [size_mib] [frame_size] [steps] [tag]   -> one JSON line: ms/step, GB/s, per-kernel ms"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
import zra_b200  # noqa: E402

size_mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
frame = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
tag = sys.argv[4] if len(sys.argv) > 4 else ""
data, archive = bench.build_archive(size_mib << 20, frame, 3, seed=7)
ctx = zra_b200.CudaContext(0)
d_in = torch.zeros(archive.size + 64, dtype=torch.uint8, device="cuda")
d_in[: archive.size] = torch.from_numpy(archive).cuda()
d_out = torch.empty(data.size, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()


def step():
    ctx.decompress_buffer(d_in.data_ptr(), archive.size, d_out.data_ptr(), data.size, st.cuda_stream)


for _ in range(3):
    step()
torch.cuda.synchronize()
ok = bool(torch.equal(d_out, torch.from_numpy(data).cuda()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(steps):
    step()
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
ctx.set_profiling(True)
for _ in range(steps):
    step()
torch.cuda.synchronize()
prof = {k: round(v[0] / steps, 4) for k, v in ctx.kernel_profile().items() if v[1]}
env = {k: v for k, v in os.environ.items() if k.startswith("ZRA_B200_")}
print(json.dumps({"tag": tag, "ok": ok, "ms_per_step": round(ms, 4), "GBps": round(data.size / ms / 1e6, 2), "env": env,
                  "kernels_ms": prof}), flush=True)
