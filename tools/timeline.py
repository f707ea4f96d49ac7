"""Dumps the chunk pipeline's event timeline (ZRA_B200_TIMELINE=1) for one decode of the bench archive: TL lines on stderr.
usage: ZRA_B200_TIMELINE=1 timeline.py [size_mib] [frame_size]"""
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
data, archive = bench.build_archive(size_mib << 20, frame, 3, seed=7)
ctx = zra_b200.CudaContext(0)
d_in = torch.zeros(archive.size + 64, dtype=torch.uint8, device="cuda")
d_in[: archive.size] = torch.from_numpy(archive).cuda()
d_out = torch.empty(data.size, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
for _ in range(3):
    ctx.decompress_buffer(d_in.data_ptr(), archive.size, d_out.data_ptr(), data.size, st.cuda_stream)
torch.cuda.synchronize()
ctx.set_profiling(True)
sys.stderr.write("TL-BEGIN\n")
ctx.decompress_buffer(d_in.data_ptr(), archive.size, d_out.data_ptr(), data.size, st.cuda_stream)
torch.cuda.synchronize()
sys.stderr.write("TL-END\n")
