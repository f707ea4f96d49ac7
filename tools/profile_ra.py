"""Small driver for ncu captures of the batched random-access path: N batches of 4 KiB reads (one per frame on average)
into a 16 KiB-frame archive (the bench's random-access leg, no timing)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import zra_b200  # noqa: E402

size_mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fs, rsz = 16384, 4096
size = size_mib << 20
data, archive = bench.build_archive(size, fs, 3, seed=107)
count = size // fs
offs = np.random.default_rng(42).integers(0, size - rsz - 1, count).astype(np.uint64)
ctx = zra_b200.CudaContext(0)
d_in = torch.zeros(archive.size + 64, dtype=torch.uint8, device="cuda")
d_in[: archive.size] = torch.from_numpy(archive).cuda()
d_off = torch.from_numpy(offs.view(np.int64)).cuda()
d_out = torch.empty(count * rsz, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    ctx.decompress_ra_batch(d_in.data_ptr(), archive.size, d_off.data_ptr(), count, d_out.data_ptr(), uniform_size=rsz,
                            stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
got = d_out.cpu().numpy().reshape(count, rsz)
print("ok", all(np.array_equal(got[i], data[int(offs[i]): int(offs[i]) + rsz]) for i in range(0, count, 997)))
