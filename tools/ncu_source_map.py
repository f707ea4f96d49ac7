"""Per-source-line instruction and stall-sample shares of one kernel from an `ncu --set full --import-source on` report.

ncu's CSV source page lists SASS only; this joins it, instruction by instruction, with `nvdisasm -g` line information of the
SAME build (the object file the .so was linked from), and prints the hottest source lines.
usage: ncu_source_map.py <report.ncu-rep> <kernel-name-substring> <object.o> [top=40]
e.g.:  ncu_source_map.py gpurun_out/r02e_full.ncu-rep k_seq_execute zra_b200/csrc/build/decode_kernels.o"""
import csv
import io
import linecache
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, kern, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src_dir = os.path.dirname(os.path.dirname(os.path.abspath(obj)))

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hi]
ie, smp, src = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
sass = []
for r in rows[hi + 1:]:
    try:
        sass.append((r[src].strip(), int(r[ie]), int(r[smp])))
    except (ValueError, IndexError):
        if sass:
            break  # a second kernel table follows the first

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    lines = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith("\t.section\t.text.") and kern in l][0]
cur, seq = None, []
for l in lines[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l):
        seq.append(cur)
if len(sass) != len(seq):
    print(f"warning: {len(sass)} profiled instructions vs {len(seq)} disassembled — is {obj} the profiled build?", file=sys.stderr)
agg = defaultdict(lambda: [0, 0])
for (_, i, s), ln in zip(sass, seq):
    agg[ln][0] += i
    agg[ln][1] += s
ti, ts = sum(v[0] for v in agg.values()), max(1, sum(v[1] for v in agg.values()))
print(f"{kern}: {ti} warp-instructions, {ts} stall samples")
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = linecache.getline(ln[0] if os.path.exists(ln[0]) else os.path.join(src_dir, os.path.basename(ln[0])), ln[1]).strip()[:100] if ln else ""
    print(f"{v[0] / ti * 100:5.1f}% inst {v[1] / ts * 100:5.1f}% smp  {os.path.basename(ln[0]) if ln else '?'}:{ln[1] if ln else 0}: {text}")
