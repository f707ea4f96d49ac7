"""Generates the committed fixtures under tests/golden/ (run HERE, where /root/reference exists).

  * <name>.zra ............ archives written by the UNMODIFIED reference (serial
                            zra::CompressBuffer via oracle/_ref/libzra_ref.so) from deterministic
                            synthetic inputs (zra_b200/synth.py); manifest.json records the recipe
                            and the SHA-256 of the plaintext.
  * dc/zNNNNNN.zst ........ frames from zstd's generative conformance tool tests/decodecorpus.c
                            (oracle/_ref/decodecorpus, seed 2024) with plaintext SHA-256 + size.
  * rle-first-block.zst ... zstd's own golden decompression file, copied as a test VECTOR
                            (zstd/tests/golden-decompression/, 45 bytes).
Usage: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refzra import ref_compress  # noqa: E402
from zra_b200 import synth  # noqa: E402


def make_input(kind, n):
    if kind == "text":
        return synth.text(n, threads=1)
    if kind == "random":
        return synth.random_bytes(n, threads=1)
    if kind == "mixed":
        return synth.mixed(n, period=16384, threads=1)
    if kind == "zeros":
        return np.zeros(n, np.uint8)
    raise ValueError(kind)


ARCHIVES = [
    # name, kind, bytes, frameSize, level, checksum
    ("text_f16384_l3", "text", 200_000, 16384, 3, True),
    ("text_f65536_l1_nock", "text", 300_000, 65536, 1, False),
    ("text_f65536_l3", "text", 262_144, 65536, 3, True),
    ("text_f262144_l3", "text", 700_000, 262144, 3, True),
    ("text_f262144_l2", "text", 524_288, 262144, 2, True),
    ("text_f4096_l19", "text", 50_000, 4096, 19, True),
    ("text_f1000_l5", "text", 10_001, 1000, 5, True),
    ("mixed_f16384_l3", "mixed", 180_000, 16384, 3, True),
    ("random_f65536_l3", "random", 150_000, 65536, 3, True),
    ("zeros_f65536_l3", "zeros", 1 << 20, 65536, 3, True),
    ("zeros_f262144_l3", "zeros", 1 << 20, 262144, 3, True),
    ("empty_f16384_l3", "zeros", 0, 16384, 3, True),
    ("tiny_f16384_l3", "text", 5, 16384, 3, True),
    ("onebyte_f16384_l3", "text", 1, 16384, 3, False),
]


def main():
    manifest = {"archives": [], "decodecorpus": []}
    for name, kind, n, fs, lvl, ck in ARCHIVES:
        data = make_input(kind, n)
        z = ref_compress(data, lvl, fs, ck)
        z.tofile(os.path.join(HERE, name + ".zra"))
        manifest["archives"].append({"name": name, "kind": kind, "bytes": n, "frameSize": fs, "level": lvl, "checksum": ck,
                                     "archiveBytes": int(z.size), "sha256": hashlib.sha256(data.tobytes()).hexdigest()})
        print(name, n, "->", z.size)
    # decodecorpus
    dc = os.path.join(HERE, "dc")
    shutil.rmtree(dc, ignore_errors=True)
    os.makedirs(dc)
    tool = os.path.join(ROOT, "oracle", "_ref", "decodecorpus")
    with tempfile.TemporaryDirectory() as tmp:
        zdir, odir = os.path.join(tmp, "z"), os.path.join(tmp, "o")
        os.makedirs(zdir); os.makedirs(odir)
        subprocess.run([tool, "-p" + zdir, "-o" + odir, "-n400", "-s2024", "--max-content-size-log=17"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        picked, total = 0, 0
        for f in sorted(os.listdir(zdir)):
            z = open(os.path.join(zdir, f), "rb").read()
            o = open(os.path.join(odir, f[:-4]), "rb").read()
            if len(z) > 24_000 or total + len(z) > 700_000:
                continue
            shutil.copy(os.path.join(zdir, f), os.path.join(dc, f))
            manifest["decodecorpus"].append({"name": f, "bytes": len(o), "sha256": hashlib.sha256(o).hexdigest()})
            picked += 1; total += len(z)
        print("decodecorpus frames:", picked, "bytes:", total)
    shutil.copy("/root/reference/submodule/zstd/tests/golden-decompression/rle-first-block.zst", os.path.join(HERE, "rle-first-block.zst"))
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
