"""Streaming zra::Compressor timing: python tools/time_compressor.py <MiB> <piece MiB> [tag]"""
import argparse, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench, zra_b200
from zra_b200 import synth
size = int(sys.argv[1]) << 20
piece = int(sys.argv[2]) << 20
data = synth.text(size, seed=207)
args = argparse.Namespace(steps=5)
orig = bench.stream_compressor
src = open(bench.__file__).read()
res = bench.stream_compressor(args, torch, zra_b200.lib(), data, size) if piece == (64 << 20) else None
if res is None:
    import types
    code = src.replace("piece = 64 << 20", f"piece = {piece}")
    mod = types.ModuleType("b2"); mod.__file__ = bench.__file__
    exec(compile(code, bench.__file__, "exec"), mod.__dict__)
    res = mod.stream_compressor(args, torch, zra_b200.lib(), data, size)
res["tag"] = sys.argv[3] if len(sys.argv) > 3 else ""
print(json.dumps(res))
