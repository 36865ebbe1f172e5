"""Loader of bindings/c/stream_pump.c: the reference's CompressionStream / DecompressionStream driver loop
(src/mod/streams.ts:78-93,139-170) in C, so that the stream API is measured without 30-50 us of Python per call.
Used by bench.py (extra.stream_api), tools/streamprof.py and tests/test_zlib_api_gpu.py."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "bindings", "c", "stream_pump.c")
OUT = os.path.join(ROOT, "bindings", "c", "_build", "libstreampump.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-Wall", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), SRC, "-o", OUT], check=True)
    return OUT


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.zs_stream_pump.restype = C.c_int
        lib.zs_stream_pump.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64,
                                       C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib = lib
    return _lib


def pump(step_fn, zs, src, dst, in_slice=32 * 1024, out_slice=64 * 1024):
    """Run numpy uint8 `src` through the stream `zs` (a capi.ZStream, initialised) with `step_fn`
    (lib.zs_stream_deflate / lib.zs_stream_inflate), output into numpy uint8 `dst`.
    Returns (rc, bytes produced, calls made)."""
    import numpy as np
    lib = load()
    obuf = np.empty(out_slice, dtype=np.uint8)
    produced, calls = C.c_uint64(0), C.c_uint64(0)
    rc = lib.zs_stream_pump(C.cast(step_fn, C.c_void_p), C.addressof(zs), src.ctypes.data, src.size, in_slice, dst.ctypes.data, dst.size,
                            obuf.ctypes.data, out_slice, C.byref(produced), C.byref(calls))
    return rc, produced.value, calls.value
