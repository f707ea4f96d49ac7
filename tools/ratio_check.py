"""Archive size of the GPU encoder against the reference's on zstd's own synthetic generator (oracle/_ref/datagen -P<n>),
text and mixed data: one JSON line per (data, frame size, level). usage: ratio_check.py [mib] [frame sizes, comma separated]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import refzra  # noqa: E402
import zra_b200  # noqa: E402
from zra_b200 import synth  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = mib << 20
gen = os.path.join(ROOT, "oracle", "_ref", "datagen")
sets = {}
for p in (20, 50, 80):
    raw = subprocess.run([gen, f"-g{n}", f"-P{p}", "-s1"], capture_output=True).stdout
    sets[f"datagen_P{p}"] = np.frombuffer(raw, dtype=np.uint8)[:n].copy()
sets["text"] = synth.text(n, seed=3)
sets["mixed"] = synth.mixed(n, period=65536, seed=3)
for name, data in sets.items():
    for fs in ((16384, 65536) if len(sys.argv) < 3 else tuple(int(x) for x in sys.argv[2].split(","))):
        for lvl in (1, 2, 3):
            ref = refzra.ref_compress_mt(data, lvl, fs, True).size
            z = zra_b200.CompressBuffer(data, lvl, fs, True)
            ok = bool(np.array_equal(refzra.ref_decompress(z), data))
            print(json.dumps({"data": name, "frame": fs, "level": lvl, "ref": int(ref), "gpu": int(z.size),
                              "delta": round(z.size / ref - 1, 4), "roundtrip_through_reference": ok}), flush=True)
