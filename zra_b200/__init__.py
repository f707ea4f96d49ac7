"""zra-b200: B200-native implementation of ZRA's frame-parallel hot path.

The product is the C-ABI shared library ``zra_b200/libzra_b200.so`` (CUDA kernels for sm_100a,
the ``Zra*`` functions of ``include/zra.h`` and the device-pointer entry points of
``include/zra_b200.h``). This Python package is only the thin ctypes binding used by the tests
and by ``bench.py``; it mirrors the reference's operator names (``CompressBuffer``,
``DecompressBuffer``, ``DecompressRA``, ``Compressor``, ``Decompressor``, ``FullDecompressor``,
``Header``) one to one. There is no CPU fallback: every call goes to the GPU library and raises
if it is missing or no CUDA device is usable.
"""
from .binding import (  # noqa: F401
    ZraError,
    StatusCode,
    lib,
    lib_path,
    GetVersion,
    GetOutputBufferSize,
    CompressBuffer,
    DecompressBuffer,
    DecompressRA,
    Header,
    Compressor,
    Decompressor,
    FullDecompressor,
    CudaContext,
    VerifyHeaderCrc,
)
