"""Shared helpers for the test-suite."""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
MANIFEST = json.load(open(os.path.join(GOLDEN, "manifest.json")))


def golden_archives():
    return [a["name"] for a in MANIFEST["archives"]]


def golden_archive(name):
    meta = next(a for a in MANIFEST["archives"] if a["name"] == name)
    return np.fromfile(os.path.join(GOLDEN, name + ".zra"), dtype=np.uint8), meta


def golden_frames():
    return [f["name"] for f in MANIFEST["decodecorpus"]]


def golden_frame(name):
    meta = next(f for f in MANIFEST["decodecorpus"] if f["name"] == name)
    return np.fromfile(os.path.join(GOLDEN, "dc", name), dtype=np.uint8), meta


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def parse_header(archive):
    """Python-side reading of the 38-byte fixed header (tests only)."""
    import struct

    frameId, headerSize, magic, version, crc, usize, table, fsize, msize = struct.unpack("<IIIHIQIII", archive[:38].tobytes())
    return dict(frameId=frameId, headerSize=headerSize, magic=magic, version=version, hash=crc, uncompressedSize=usize,
                tableSize=table, frameSize=fsize, metaSize=msize, size=headerSize + 8)


def seek_table(archive):
    h = parse_header(archive)
    t = archive[38 + h["metaSize"]: 38 + h["metaSize"] + 5 * h["tableSize"]].reshape(-1, 5).astype(np.uint64)
    return t[:, 0] | (t[:, 1] << 8) | (t[:, 2] << 16) | (t[:, 3] << 24) | (t[:, 4] << 32)
