"""Prints the hot SASS lines of one kernel from `ncu --page source --csv` output (stall samples, executions)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minex = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, isrc, iss, iex, ith = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"))
data = []
for r in rows[h + 1:]:
    if len(r) <= ith or not r[ia].startswith("0x"):
        continue
    data.append((int(r[ia], 16), r[isrc].strip(), int(r[iss]), int(r[iex]), r[ith]))
base = data[0][0]
tot = sum(d[2] for d in data)
totex = sum(d[3] for d in data)
print("total samples", tot, "total inst", totex)
for a, s, n, e, t in data:
    if e >= minex:
        print(f"{a - base:5x} {n:6d} {100 * n / max(tot, 1):5.1f}% ex={e:9d} thr={t:>5} {s[:80]}")
