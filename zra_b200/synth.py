"""Deterministic synthetic inputs: Zipf-word text, PRNG bytes, 50/50 mixed."""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
VOCAB_SEED = 7
RANDOM_SEED = 11
CHUNK = 4 << 20


def _l():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libzra_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run __graft_entry__.build()")
        _lib = C.CDLL(path)
        _lib.zra_synth_text.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64]
        _lib.zra_synth_random.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
    return _lib


def _fill(out, kind, seed, threads):
    n = out.size
    lib = _l()

    def job(i):
        a = i * CHUNK
        m = min(CHUNK, n - a)
        p = out.ctypes.data + a
        if kind == "text":
            lib.zra_synth_text(p, m, VOCAB_SEED, seed * 1000003 + i)
        else:
            lib.zra_synth_random(p, m, seed * 1000003 + i)

    chunks = (n + CHUNK - 1) // CHUNK
    if threads <= 1 or chunks <= 1:
        for i in range(chunks):
            job(i)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(job, range(chunks)))
    return out


def text(n, seed=VOCAB_SEED, threads=None, out=None):
    """n bytes of Zipf-word text; independent 4 MiB chunks share one vocabulary."""
    out = np.empty(n, np.uint8) if out is None else out
    return _fill(out, "text", seed, threads or min(32, os.cpu_count() or 1))


def random_bytes(n, seed=RANDOM_SEED, threads=None, out=None):
    out = np.empty(n, np.uint8) if out is None else out
    return _fill(out, "random", seed, threads or min(32, os.cpu_count() or 1))


def mixed(n, period=65536, seed=VOCAB_SEED, threads=None):
    """Alternating `period`-byte runs of text and incompressible bytes (BASELINE config 3)."""
    out = text(n, seed, threads)
    rnd = random_bytes((n + 1) // 2 + period, RANDOM_SEED, threads)
    k = 0
    for a in range(period, n, 2 * period):
        m = min(period, n - a)
        out[a:a + m] = rnd[k:k + m]
        k += m
    return out
