"""Condenses gpurun_out/<tag>_launches.csv (ncu launch list) and <tag>_full.ncu-rep (ncu --set full) into the
small tracked summaries under profiles/: per-kernel launch counts / time share, and the key full-set metrics."""
import csv
import io
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_tag = sys.argv[2] if len(sys.argv) > 2 else tag
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
] + ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s for s in (
    "long_scoreboard", "short_scoreboard", "wait", "branch_resolving", "barrier", "math_pipe_throttle", "mio_throttle",
    "lg_throttle", "no_instruction", "not_selected", "dispatch_stall")]

lf = os.path.join(G, tag + "_launches.csv")
if os.path.exists(lf):
    lines = [l for l in open(lf) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("zrab::", "")
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e6
    total = sum(a[1] for n, a in agg.items() if n.startswith("k_"))
    with open(os.path.join(P, out_tag + "_launches_summary.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ra\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("kernel,launches,total_ms,share_of_zra_kernels,first_grid,first_block\n")
        for n, a in agg.items():
            share = a[1] / total if n.startswith("k_") and total else 0
            f.write(f"{n},{a[0]},{a[1]:.4f},{share:.4f},\"{a[2]}\",\"{a[3]}\"\n")
    print("wrote launches summary:", len(rows), "launches")

for suffix in ("_full", "_enc_full", "_ra_full"):
  rep = os.path.join(G, tag + suffix + ".ncu-rep")
  rawcsv = os.path.join(G, tag + suffix + "_raw.csv")
  if os.path.exists(rep) or os.path.exists(rawcsv):
      raw = (subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
             if os.path.exists(rep) else open(rawcsv).read())
      rows = list(csv.reader(io.StringIO(raw)))
      hdr, units = rows[0], rows[1]
      kn = hdr.index("Kernel Name")
      with open(os.path.join(P, out_tag + "_ncu" + suffix + "_summary.csv"), "w") as f:
          names = [r[kn].split("(")[0].replace(", ", "_").replace(",", "_") for r in rows[2:]]
          sys.path.insert(0, ROOT)
          import bench  # the stamp bench.py checks before it trusts a capture's DRAM traffic
          f.write("# source_sha256=" + bench.kernel_source_hash() + "\n")
          f.write("metric,unit," + ",".join(names) + "\n")
          for k in KEYS:
              if k in hdr:
                  i = hdr.index(k)
                  f.write(k + "," + units[i] + "," + ",".join('"%s"' % r[i] if "," in r[i] else r[i] for r in rows[2:]) + "\n")
      print("wrote full summary:", len(rows) - 2, "kernels")
