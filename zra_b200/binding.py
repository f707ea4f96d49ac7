"""ctypes binding of libzra_b200.so — mirrors the reference's API names (zra.h / zra.hpp).

Reference surface being mirrored: /root/reference/include/zra.h:56-263 (C) and
include/zra.hpp:88-322 (C++). Buffers are ``bytes`` / ``bytearray`` / numpy uint8 arrays on the
host; ``CudaContext`` exposes the additive device-pointer entry points of include/zra_b200.h.
"""
import ctypes as C
import enum
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "libzra_b200.so")


class StatusCode(enum.IntEnum):
    Success = 0
    ZStdError = 1
    ZraVersionLow = 2
    HeaderInvalid = 3
    HeaderIncomplete = 4
    OutOfBoundsAccess = 5
    OutputBufferTooSmall = 6
    CompressedSizeTooLarge = 7
    InputFrameSizeMismatch = 8


class ZraStatus(C.Structure):
    _fields_ = [("zra", C.c_int), ("zstd", C.c_int)]


class ZraError(Exception):
    """Python twin of zra::Exception: ``code`` (StatusCode) and ``zstd_code`` (ZSTD_ErrorCode)."""

    def __init__(self, code, zstd_code=0, message=None):
        self.code = StatusCode(code)
        self.zstd_code = int(zstd_code)
        super().__init__(message or f"{self.code.name} (zstd={self.zstd_code})")


class CudaFrame(C.Structure):
    _fields_ = [("srcOffset", C.c_uint64), ("dstOffset", C.c_uint64), ("srcSize", C.c_uint32),
                ("dstCapacity", C.c_uint32), ("exact", C.c_uint32), ("reserved", C.c_uint32)]


READ_FN = C.CFUNCTYPE(None, C.c_size_t, C.c_size_t, C.c_void_p)

_lib = None


def lib():
    """Loads the CUDA library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(zra-b200 has no CPU fallback)")
    L = C.CDLL(path)
    vp, sz, u8, u32, u64 = C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint32, C.c_uint64
    P = C.POINTER
    sig = {
        "ZraGetVersion": (C.c_uint16, []),
        "ZraGetErrorString": (C.c_char_p, [ZraStatus]),
        "ZraCreateHeader": (ZraStatus, [P(vp), READ_FN]),
        "ZraCreateHeader2": (ZraStatus, [P(vp), vp, sz]),
        "ZraDeleteHeader": (None, [vp]),
        "ZraGetVersionWithHeader": (sz, [vp]),
        "ZraGetHeaderSizeWithHeader": (sz, [vp]),
        "ZraGetUncompressedSizeWithHeader": (sz, [vp]),
        "ZraGetFrameSizeWithHeader": (sz, [vp]),
        "ZraGetMetadataSize": (sz, [vp]),
        "ZraGetMetadata": (None, [vp, vp]),
        "ZraGetCompressedOutputBufferSize": (sz, [sz, sz]),
        "ZraCompressBuffer": (ZraStatus, [vp, sz, vp, P(sz), C.c_int8, u32, C.c_bool, vp, sz]),
        "ZraDecompressBuffer": (ZraStatus, [vp, sz, vp]),
        "ZraDecompressRA": (ZraStatus, [vp, sz, vp, sz, sz]),
        "ZraCreateCompressor": (ZraStatus, [P(vp), sz, C.c_int8, u32, C.c_bool, vp, sz]),
        "ZraDeleteCompressor": (None, [vp]),
        "ZraGetOutputBufferSizeWithCompressor": (sz, [vp, sz]),
        "ZraCompressWithCompressor": (ZraStatus, [vp, vp, sz, vp, P(sz)]),
        "ZraGetHeaderSizeWithCompressor": (sz, [vp]),
        "ZraGetHeaderWithCompressor": (ZraStatus, [vp, vp]),
        "ZraCreateDecompressor": (ZraStatus, [P(vp), READ_FN, sz]),
        "ZraDeleteDecompressor": (None, [vp]),
        "ZraGetHeaderWithDecompressor": (vp, [vp]),
        "ZraDecompressWithDecompressor": (ZraStatus, [vp, sz, sz, vp]),
        "ZraCreateFullDecompressor": (ZraStatus, [P(vp), READ_FN, sz]),
        "ZraDeleteFullDecompressor": (None, [vp]),
        "ZraGetHeaderWithFullDecompressor": (vp, [vp]),
        "ZraDecompressWithFullDecompressor": (ZraStatus, [vp, vp, sz, P(sz)]),
        "ZraCudaCreateContext": (ZraStatus, [P(vp), C.c_int]),
        "ZraCudaDestroyContext": (None, [vp]),
        "ZraCudaGetLastError": (C.c_char_p, [vp]),
        "ZraCudaGetLaunchCount": (u64, [vp]),
        "ZraCudaSetProfiling": (None, [vp, C.c_int]),
        "ZraCudaGetKernelProfile": (C.c_int, [vp, C.c_int, P(C.c_char_p), P(C.c_double), P(u64)]),
        "ZraCudaDecodeFrames": (ZraStatus, [vp, vp, sz, P(CudaFrame), u32, vp, P(u32), P(u32), vp]),
        "ZraCudaDecompressBuffer": (ZraStatus, [vp, vp, sz, vp, sz, vp]),
        "ZraCudaDecompressFrames": (ZraStatus, [vp, vp, sz, u64, u64, vp, sz, vp]),
        "ZraCudaCompressBuffer": (ZraStatus, [vp, vp, sz, vp, sz, P(sz), C.c_int8, u32, C.c_bool, vp, sz, vp]),
        "ZraCudaCompressFrames": (ZraStatus, [vp, vp, sz, u32, C.c_int8, C.c_bool, vp, sz, P(u64), P(sz), vp]),
        "ZraVerifyHeaderCrc": (ZraStatus, [vp, sz]),
        "ZraShardHeaderSize": (sz, [u64, sz]),
        "ZraShardBuildHeader": (ZraStatus, [u64, u32, vp, sz, P(u64), u64, vp, sz]),
        "ZraCudaDecompressRABatch": (ZraStatus, [vp, vp, sz, vp, vp, vp, u32, u32, u64, vp, P(u64), P(u64), vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(status):
    if status.zra != 0:
        msg = lib().ZraGetErrorString(status)
        raise ZraError(status.zra, status.zstd, msg.decode() if msg else None)


def _as_array(buf):
    """A uint8 numpy view of a bytes-like object (no copy when possible)."""
    if isinstance(buf, np.ndarray):
        a = buf if buf.dtype == np.uint8 else buf.view(np.uint8)
        return np.ascontiguousarray(a).reshape(-1)
    return np.frombuffer(buf, dtype=np.uint8)


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a.size else C.c_void_p(0)


# ------------------------------------------------------------------ free functions
def GetVersion():
    return lib().ZraGetVersion()


def GetOutputBufferSize(inputSize, frameSize, metaSize=0):
    # zra.h only exposes the no-metadata form; the C++ form adds metaSize (source/zra.cpp:189-192)
    return lib().ZraGetCompressedOutputBufferSize(inputSize, frameSize) + metaSize


def CompressBuffer(data, compressionLevel=0, frameSize=16384, checksum=True, meta=b""):
    src = _as_array(data)
    m = _as_array(meta)
    out = np.empty(lib().ZraGetCompressedOutputBufferSize(src.size, frameSize), np.uint8)
    n = C.c_size_t(0)
    _check(lib().ZraCompressBuffer(_ptr(src), src.size, _ptr(out), C.byref(n), compressionLevel, frameSize, checksum,
                                   _ptr(m), m.size))
    return out[: n.value]


def DecompressBuffer(archive, out=None):
    src = _as_array(archive)
    if src.size < 26:
        raise ZraError(StatusCode.OutOfBoundsAccess)
    size = int(np.frombuffer(src[18:26].tobytes(), "<u8")[0])
    if out is None:
        out = np.empty(size, np.uint8)
    _check(lib().ZraDecompressBuffer(_ptr(src), src.size, _ptr(out)))
    return out


def DecompressRA(archive, offset, size, out=None):
    src = _as_array(archive)
    if out is None:
        out = np.empty(size, np.uint8)
    _check(lib().ZraDecompressRA(_ptr(src), src.size, _ptr(out), offset, size))
    return out


# ------------------------------------------------------------------ Header
class Header:
    def __init__(self, source=None, _borrowed=None, _owner=None):
        self._own = _borrowed is None
        self._keep = _owner
        if _borrowed is not None:
            self._h = C.c_void_p(_borrowed)
            return
        h = C.c_void_p()
        if callable(source):
            self._cb = _wrap_reader(source)
            _check(lib().ZraCreateHeader(C.byref(h), self._cb))
        else:
            self._buf = _as_array(source)
            _check(lib().ZraCreateHeader2(C.byref(h), _ptr(self._buf), self._buf.size))
        self._h = h

    version = property(lambda s: lib().ZraGetVersionWithHeader(s._h))
    size = property(lambda s: lib().ZraGetHeaderSizeWithHeader(s._h))
    uncompressedSize = property(lambda s: lib().ZraGetUncompressedSizeWithHeader(s._h))
    frameSize = property(lambda s: lib().ZraGetFrameSizeWithHeader(s._h))
    metaSize = property(lambda s: lib().ZraGetMetadataSize(s._h))

    def GetMetadata(self):
        out = np.empty(self.metaSize, np.uint8)
        lib().ZraGetMetadata(self._h, _ptr(out))
        return out.tobytes()

    def __del__(self):
        if getattr(self, "_own", False) and getattr(self, "_h", None):
            lib().ZraDeleteHeader(self._h)
            self._h = None


def VerifyHeaderCrc(archive):
    """include/zra_b200.h: ZraVerifyHeaderCrc — recomputes the header CRC-32 the reference stores but never checks.
    Returns True / False; raises for a truncated buffer."""
    a = _as_array(archive)
    st = lib().ZraVerifyHeaderCrc(_ptr(a), a.size)
    if st.zra == StatusCode.HeaderInvalid:
        return False
    _check(st)
    return True


def _wrap_reader(fn):
    """fn(offset, size) -> bytes-like, adapted to the C read callback."""

    def cb(offset, size, buffer):
        data = fn(offset, size)
        if size:
            C.memmove(buffer, bytes(data) if not isinstance(data, (bytes, bytearray)) else data, size)

    return READ_FN(cb)


# ------------------------------------------------------------------ Compressor
class Compressor:
    def __init__(self, size, compressionLevel=0, frameSize=16384, checksum=True, meta=b""):
        m = _as_array(meta)
        self._c = C.c_void_p()
        _check(lib().ZraCreateCompressor(C.byref(self._c), size, compressionLevel, frameSize, checksum, _ptr(m), m.size))

    def GetOutputBufferSize(self, inputSize):
        return lib().ZraGetOutputBufferSizeWithCompressor(self._c, inputSize)

    def Compress(self, data):
        src = _as_array(data)
        out = np.empty(self.GetOutputBufferSize(src.size), np.uint8)
        n = C.c_size_t(0)
        _check(lib().ZraCompressWithCompressor(self._c, _ptr(src), src.size, _ptr(out), C.byref(n)))
        return out[: n.value]

    def GetHeaderSize(self):
        return lib().ZraGetHeaderSizeWithCompressor(self._c)

    def GetHeader(self):
        out = np.empty(self.GetHeaderSize(), np.uint8)
        _check(lib().ZraGetHeaderWithCompressor(self._c, _ptr(out)))
        return out

    def __del__(self):
        if getattr(self, "_c", None):
            lib().ZraDeleteCompressor(self._c)
            self._c = None


# ------------------------------------------------------------------ Decompressor / FullDecompressor
class Decompressor:
    def __init__(self, readFunction, maxCacheSize=20 * 1024 * 1024):
        self._cb = _wrap_reader(readFunction)
        self._d = C.c_void_p()
        _check(lib().ZraCreateDecompressor(C.byref(self._d), self._cb, maxCacheSize))
        self.header = Header(_borrowed=lib().ZraGetHeaderWithDecompressor(self._d), _owner=self)

    def Decompress(self, offset, size, out=None):
        if out is None:
            out = np.empty(size, np.uint8)
        _check(lib().ZraDecompressWithDecompressor(self._d, offset, size, _ptr(out)))
        return out

    def __del__(self):
        if getattr(self, "_d", None):
            lib().ZraDeleteDecompressor(self._d)
            self._d = None


class FullDecompressor:
    def __init__(self, readFunction):
        self._cb = _wrap_reader(readFunction)
        self._d = C.c_void_p()
        _check(lib().ZraCreateFullDecompressor(C.byref(self._d), self._cb, 0))
        self.header = Header(_borrowed=lib().ZraGetHeaderWithFullDecompressor(self._d), _owner=self)

    def Decompress(self, out):
        """Fills `out` (numpy uint8) with as many frames as fit; returns the byte count (0 = end)."""
        n = C.c_size_t(0)
        _check(lib().ZraDecompressWithFullDecompressor(self._d, _ptr(out), out.size, C.byref(n)))
        return n.value

    def __del__(self):
        if getattr(self, "_d", None):
            lib().ZraDeleteFullDecompressor(self._d)
            self._d = None


# ------------------------------------------------------------------ device-pointer entry points
class CudaContext:
    """include/zra_b200.h: operations on buffers that already live in HBM (raw device pointers)."""

    def __init__(self, device=-1):
        self._c = C.c_void_p()
        st = lib().ZraCudaCreateContext(C.byref(self._c), device)
        if st.zra != 0:
            msg = lib().ZraCudaGetLastError(self._c).decode()
            lib().ZraCudaDestroyContext(self._c)
            self._c = None
            raise ZraError(st.zra, st.zstd, msg)

    def last_error(self):
        return lib().ZraCudaGetLastError(self._c).decode()

    def launch_count(self):
        return lib().ZraCudaGetLaunchCount(self._c)

    def set_profiling(self, on):
        lib().ZraCudaSetProfiling(self._c, 1 if on else 0)

    def kernel_profile(self):
        """{kernel name: (total ms, launches)} accumulated since set_profiling()."""
        out, i = {}, 1
        name, ms, n = C.c_char_p(), C.c_double(), C.c_uint64()
        while lib().ZraCudaGetKernelProfile(self._c, i, C.byref(name), C.byref(ms), C.byref(n)):
            out[name.value.decode()] = (ms.value, n.value)
            i += 1
        return out

    def _raise(self, st):
        if st.zra != 0:
            raise ZraError(st.zra, st.zstd, (lib().ZraGetErrorString(st) or b"").decode() + " | " + self.last_error())

    def decode_frames(self, d_src, src_size, frames, d_dst, stream=0, want_sizes=False):
        """frames: iterable of (srcOffset, srcSize, dstOffset, dstCapacity, exact)."""
        n = len(frames)
        arr = (CudaFrame * n)()
        for i, (so, sl, do, dc, ex) in enumerate(frames):
            arr[i] = CudaFrame(so, do, sl, dc, ex, 0)
        sizes = (C.c_uint32 * n)() if want_sizes else None
        failed = C.c_uint32(0xFFFFFFFF)
        st = lib().ZraCudaDecodeFrames(self._c, C.c_void_p(d_src), src_size, arr, n, C.c_void_p(d_dst), sizes, C.byref(failed),
                                       C.c_void_p(stream))
        if st.zra != 0:
            e = ZraError(st.zra, st.zstd, f"frame {failed.value}: zstd error {st.zstd} | {self.last_error()}")
            e.failed_frame = failed.value
            raise e
        return list(sizes) if want_sizes else None

    def decompress_buffer(self, d_archive, archive_size, d_out, out_capacity, stream=0):
        self._raise(lib().ZraCudaDecompressBuffer(self._c, C.c_void_p(d_archive), archive_size, C.c_void_p(d_out), out_capacity,
                                                  C.c_void_p(stream)))

    def decompress_frames(self, d_archive, archive_size, first_frame, frame_count, d_out, out_capacity, stream=0):
        self._raise(lib().ZraCudaDecompressFrames(self._c, C.c_void_p(d_archive), archive_size, first_frame, frame_count,
                                                  C.c_void_p(d_out), out_capacity, C.c_void_p(stream)))

    def decompress_ra_batch(self, d_archive, archive_size, d_offsets, count, d_out, uniform_size=0, d_sizes=0, d_out_offsets=0,
                            max_size=0, stream=0):
        """Batched random access; all request arrays are device pointers (0 = absent). Returns the number of
        frames decoded after de-duplication."""
        unique, bad = C.c_uint64(0), C.c_uint64(0)
        st = lib().ZraCudaDecompressRABatch(self._c, C.c_void_p(d_archive), archive_size, C.c_void_p(d_offsets),
                                            C.c_void_p(d_sizes or None), C.c_void_p(d_out_offsets or None), uniform_size, max_size,
                                            count, C.c_void_p(d_out), C.byref(unique), C.byref(bad), C.c_void_p(stream))
        if st.zra != 0:
            e = ZraError(st.zra, st.zstd, (lib().ZraGetErrorString(st) or b"").decode() + f" | request {bad.value} | " + self.last_error())
            e.bad_request = bad.value
            raise e
        return unique.value

    def compress_buffer(self, d_in, in_size, d_out, out_capacity, level=0, frame_size=16384, checksum=True, meta=b"", stream=0):
        """Device-resident CompressBuffer; returns the archive size."""
        m = _as_array(meta)
        n = C.c_size_t(0)
        self._raise(lib().ZraCudaCompressBuffer(self._c, C.c_void_p(d_in), in_size, C.c_void_p(d_out), out_capacity, C.byref(n), level,
                                                frame_size, checksum, _ptr(m), m.size, C.c_void_p(stream)))
        return n.value

    def __del__(self):
        if getattr(self, "_c", None):
            lib().ZraCudaDestroyContext(self._c)
            self._c = None
