"""ctypes access to the CHECKERS: oracle/libzra_oracle.so (C restatement) and, when it has been
built, oracle/_ref/libzra_ref.so (the unmodified reference + MT harness). Test infrastructure."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libzra_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libzra_ref.so")

_oracle = None
_ref = None


class OracleError(Exception):
    def __init__(self, zra, zstd):
        self.zra, self.zstd = zra, zstd
        super().__init__(f"oracle: zra={zra} zstd={zstd}")


def oracle():
    global _oracle
    if _oracle is None:
        L = C.CDLL(ORACLE_SO)
        vp, sz, u32, u64, i64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int64
        L.zra_oracle_crc32.argtypes = [vp, sz, u32]; L.zra_oracle_crc32.restype = u32
        L.zra_oracle_xxh64.argtypes = [vp, sz, u64]; L.zra_oracle_xxh64.restype = u64
        L.zra_oracle_zstd_decompress.argtypes = [vp, sz, vp, sz]; L.zra_oracle_zstd_decompress.restype = i64
        L.zra_oracle_compress_bound.argtypes = [u64]; L.zra_oracle_compress_bound.restype = u64
        L.zra_oracle_output_buffer_size.argtypes = [u64, u32, u32]; L.zra_oracle_output_buffer_size.restype = u64
        L.zra_oracle_header_crc.argtypes = [vp, sz]; L.zra_oracle_header_crc.restype = u32
        L.zra_oracle_build_header.argtypes = [vp, u64, u32, vp, u32, C.POINTER(u64), u32]; L.zra_oracle_build_header.restype = u64
        L.zra_oracle_decompress_buffer.argtypes = [vp, sz, vp, sz, C.POINTER(C.c_int)]; L.zra_oracle_decompress_buffer.restype = i64
        L.zra_oracle_decompress_ra.argtypes = [vp, sz, vp, sz, u64, u64, C.c_int, C.POINTER(C.c_int)]
        L.zra_oracle_decompress_ra.restype = i64
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        vp, sz, u32, u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64
        P = C.POINTER

        class St(C.Structure):
            _fields_ = [("zra", C.c_int), ("zstd", C.c_int)]

        L.St = St
        L.ref_compress_mt.argtypes = [vp, sz, vp, sz, P(sz), C.c_int, u32, C.c_int, C.c_int]
        L.ref_decompress_mt.argtypes = [vp, sz, vp, sz, C.c_int, P(C.c_int)]
        L.ref_ra_mt.argtypes = [vp, sz, P(u64), sz, sz, vp, C.c_int, P(C.c_int)]
        L.ref_ra_inmemory.argtypes = [vp, sz, P(u64), sz, sz, vp, P(C.c_int)]
        L.ref_zstd_decompress.argtypes = [vp, sz, vp, sz]; L.ref_zstd_decompress.restype = C.c_longlong
        L.ref_crc32.argtypes = [vp, sz, C.c_uint, C.c_int]; L.ref_crc32.restype = C.c_uint
        L.XXH64.argtypes = [vp, sz, u64]; L.XXH64.restype = u64
        L.ZraGetCompressedOutputBufferSize.argtypes = [sz, sz]; L.ZraGetCompressedOutputBufferSize.restype = sz
        L.ZraCompressBuffer.argtypes = [vp, sz, vp, P(sz), C.c_int8, u32, C.c_bool, vp, sz]; L.ZraCompressBuffer.restype = St
        L.ZraDecompressBuffer.argtypes = [vp, sz, vp]; L.ZraDecompressBuffer.restype = St
        L.ZraDecompressRA.argtypes = [vp, sz, vp, sz, sz]; L.ZraDecompressRA.restype = St
        L.ZSTD_compressBound.argtypes = [sz]; L.ZSTD_compressBound.restype = sz
        _ref = L
    return _ref


def _arr(b):
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b.view(np.uint8)).reshape(-1)
    return np.frombuffer(b, np.uint8)


def _p(a):
    return C.c_void_p(a.ctypes.data if a.size else 0)


# ---------------------------------------------------------------- oracle wrappers
def oracle_zstd_decompress(frame, cap):
    src = _arr(frame)
    out = np.empty(cap, np.uint8)
    r = oracle().zra_oracle_zstd_decompress(_p(out), cap, _p(src), src.size)
    if r < 0:
        raise OracleError(1, -r)
    return out[:r]


def oracle_decompress_buffer(archive):
    src = _arr(archive)
    size = int(np.frombuffer(src[18:26].tobytes(), "<u8")[0]) if src.size >= 26 else 0
    out = np.empty(size, np.uint8)
    st = (C.c_int * 2)()
    r = oracle().zra_oracle_decompress_buffer(_p(src), src.size, _p(out), size, st)
    if r < 0:
        raise OracleError(st[0], st[1])
    return out[:r]


def oracle_decompress_ra(archive, offset, size, in_memory_quirk=True):
    src = _arr(archive)
    out = np.empty(size, np.uint8)
    st = (C.c_int * 2)()
    r = oracle().zra_oracle_decompress_ra(_p(src), src.size, _p(out), size, offset, size, 1 if in_memory_quirk else 0, st)
    if r < 0:
        raise OracleError(st[0], st[1])
    return out


def oracle_crc32(data, prev=0):
    a = _arr(data)
    return oracle().zra_oracle_crc32(_p(a), a.size, prev)


def oracle_xxh64(data, seed=0):
    a = _arr(data)
    return oracle().zra_oracle_xxh64(_p(a), a.size, seed)


# ---------------------------------------------------------------- reference wrappers
def ref_compress(data, level=3, frame_size=16384, checksum=True, meta=b""):
    """The reference's own serial zra::CompressBuffer through its C API."""
    L = ref()
    src = _arr(data)
    m = _arr(meta)
    out = np.empty(L.ZraGetCompressedOutputBufferSize(src.size, frame_size) + m.size, np.uint8)
    n = C.c_size_t(0)
    st = L.ZraCompressBuffer(_p(src), src.size, _p(out), C.byref(n), level, frame_size, checksum, _p(m), m.size)
    if st.zra:
        raise OracleError(st.zra, st.zstd)
    return out[: n.value].copy()


def ref_compress_mt(data, level=3, frame_size=16384, checksum=True, threads=None):
    L = ref()
    src = _arr(data)
    out = np.empty(L.ZraGetCompressedOutputBufferSize(src.size, frame_size), np.uint8)
    n = C.c_size_t(0)
    rc = L.ref_compress_mt(_p(src), src.size, _p(out), out.size, C.byref(n), level, frame_size, 1 if checksum else 0,
                           threads or (os.cpu_count() or 1))
    if rc:
        raise OracleError(rc, 0)
    return out[: n.value]


def ref_decompress(archive):
    L = ref()
    src = _arr(archive)
    size = int(np.frombuffer(src[18:26].tobytes(), "<u8")[0])
    out = np.empty(size, np.uint8)
    st = L.ZraDecompressBuffer(_p(src), src.size, _p(out))
    if st.zra:
        raise OracleError(st.zra, st.zstd)
    return out


def ref_decompress_ra(archive, offset, size):
    L = ref()
    src = _arr(archive)
    out = np.empty(size, np.uint8)
    st = L.ZraDecompressRA(_p(src), src.size, _p(out), offset, size)
    if st.zra:
        raise OracleError(st.zra, st.zstd)
    return out


def ref_zstd_decompress(frame, cap):
    src = _arr(frame)
    out = np.empty(cap, np.uint8)
    r = ref().ref_zstd_decompress(_p(out), cap, _p(src), src.size)
    if r < 0:
        raise OracleError(1, -r)
    return out[:r]


def system_zstd_decompress(data, cap):
    """Stock libzstd 1.5.5 of the image: the 'any stock zstd decoder' check."""
    L = C.CDLL("libzstd.so.1")
    L.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.ZSTD_decompress.restype = C.c_size_t
    L.ZSTD_isError.argtypes = [C.c_size_t]
    src = _arr(data)
    out = np.empty(max(cap, 1), np.uint8)
    r = L.ZSTD_decompress(_p(out), cap, _p(src), src.size)
    if L.ZSTD_isError(r):
        raise OracleError(1, int((1 << 64) - r))
    return out[:r]


class _ZBuf(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]


def system_zstd_decompress_stream(src, cap, feed=1 << 16):
    """Stock libzstd 1.5.5 through its STREAMING decoder (ZSTD_decompressStream), which sizes its window and block
    buffers from each frame header's Window_Descriptor and refuses offsets beyond it — unlike the one-shot call."""
    L = C.CDLL("libzstd.so.1")
    L.ZSTD_createDStream.restype = C.c_void_p
    L.ZSTD_freeDStream.argtypes = [C.c_void_p]
    L.ZSTD_decompressStream.argtypes = [C.c_void_p, C.POINTER(_ZBuf), C.POINTER(_ZBuf)]
    L.ZSTD_decompressStream.restype = C.c_size_t
    L.ZSTD_isError.argtypes = [C.c_size_t]
    L.ZSTD_getErrorName.argtypes = [C.c_size_t]
    L.ZSTD_getErrorName.restype = C.c_char_p
    src = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.empty(max(cap, 1), np.uint8)
    ds = L.ZSTD_createDStream()
    try:
        o = _ZBuf(out.ctypes.data, cap, 0)
        at = 0
        while at < src.size:
            n = min(feed, src.size - at)
            i = _ZBuf(src.ctypes.data + at, n, 0)
            while i.pos < i.size:
                r = L.ZSTD_decompressStream(ds, C.byref(o), C.byref(i))
                if L.ZSTD_isError(r):
                    raise RuntimeError("libzstd stream: " + L.ZSTD_getErrorName(r).decode())
            at += n
        return out[: o.pos]
    finally:
        L.ZSTD_freeDStream(ds)
