"""ctypes binding of the CPU oracle (oracle/_build/libzsoracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/zs_oracle.h.  Imported by tests/, by
``__graft_entry__.smoke()`` and by bench.py's ``cpu_baseline`` / ``--impl reference`` legs; the
product package (zlib-streams-ts_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libzsoracle.so")

Z_OK, Z_STREAM_END, Z_NEED_DICT = 0, 1, 2
Z_STREAM_ERROR, Z_DATA_ERROR, Z_MEM_ERROR, Z_BUF_ERROR = -2, -3, -4, -5
Z_NO_FLUSH, Z_SYNC_FLUSH, Z_FINISH = 0, 2, 4


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p, u64p, u32p, i32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
        L.zo_adler32.restype = C.c_uint32
        L.zo_adler32.argtypes = [C.c_uint32, u8p, C.c_size_t]
        L.zo_crc32.restype = C.c_uint32
        L.zo_crc32.argtypes = [C.c_uint32, u8p, C.c_size_t]
        L.zo_crc32_combine.restype = C.c_uint32
        L.zo_crc32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.zo_adler32_combine.restype = C.c_uint32
        L.zo_adler32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.zo_inflate_table.restype = C.c_int
        L.zo_inflate_table.argtypes = [C.c_int, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int]
        L.zo_inflate_new.restype = C.c_void_p
        L.zo_inflate_free.argtypes = [C.c_void_p]
        for name in ("zo_inflate_init2", "zo_inflate"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        for name in ("zo_inflate_reset", "zo_inflate_end", "zo_inflate_mode"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p]
        L.zo_inflate_set_dictionary.restype = C.c_int
        L.zo_inflate_set_dictionary.argtypes = [C.c_void_p, u8p, C.c_size_t]
        L.zo_inflate_set_input.argtypes = [C.c_void_p, u8p, C.c_size_t]
        L.zo_inflate_set_output.argtypes = [C.c_void_p, u8p, C.c_size_t]
        for name in ("zo_inflate_avail_in", "zo_inflate_avail_out"):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("zo_inflate_total_in", "zo_inflate_total_out"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p]
        L.zo_inflate_adler.restype = C.c_uint32
        L.zo_inflate_adler.argtypes = [C.c_void_p]
        L.zo_inflate_msg.restype = C.c_char_p
        L.zo_inflate_msg.argtypes = [C.c_void_p]
        L.zo_inflate_oneshot.restype = C.c_int
        L.zo_inflate_oneshot.argtypes = [u8p, C.c_size_t, C.c_int, u8p, C.c_size_t, u8p, C.c_size_t,
                                         C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]
        L.zo_deflate_bound.restype = C.c_size_t
        L.zo_deflate_bound.argtypes = [C.c_size_t, C.c_int]
        L.zo_deflate_oneshot.restype = C.c_int64
        L.zo_deflate_oneshot.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, u8p, C.c_size_t, C.c_int, u8p,
                                         C.c_size_t]
        L.zo_deflate_oneshot2.restype = C.c_int64
        L.zo_deflate_oneshot2.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_size_t,
                                          C.c_int, u8p, C.c_size_t]
        L.zo_build_tree.restype = C.c_int
        L.zo_build_tree.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_uint32)]
        L.zo_max_threads.restype = C.c_int
        L.zo_deflate_chunks_mt.restype = C.c_int64
        L.zo_deflate_chunks_mt.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int]
        L.zo_deflate_records_mt.restype = C.c_int64
        L.zo_deflate_records_mt.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, u8p, C.c_size_t, u64p, C.c_int]
        L.zo_inflate_batch_mt.restype = C.c_int64
        L.zo_inflate_batch_mt.argtypes = [u8p, u64p, C.c_size_t, C.c_int, u8p, u64p, u64p, u32p, i32p, C.c_int]
        L.zo_checksum_mt.restype = C.c_uint32
        L.zo_checksum_mt.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_int, C.c_int]
        _lib = L
    return _lib


def _buf(data) -> tuple:
    """Return (ctypes pointer value, length, keepalive) for bytes-like data."""
    if data is None:
        return None, 0, None
    b = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    arr = (C.c_ubyte * max(len(b), 1)).from_buffer_copy(b if len(b) else b"\0")
    return C.addressof(arr), len(b), arr


def adler32(data, value: int = 1) -> int:
    p, n, k = _buf(data)
    return lib().zo_adler32(value, p, n)


def crc32(data, value: int = 0) -> int:
    p, n, k = _buf(data)
    return lib().zo_crc32(value, p, n)


def crc32_combine(c1: int, c2: int, len2: int) -> int:
    return lib().zo_crc32_combine(c1, c2, len2)


def adler32_combine(a1: int, a2: int, len2: int) -> int:
    return lib().zo_adler32_combine(a1, a2, len2)


def inflate(data, window_bits: int = 15, out_cap: int | None = None, dictionary=None):
    """One-shot inflate(strm, Z_FINISH).  Returns (ret, output bytes, bytes consumed, check)."""
    p, n, k = _buf(data)
    dp, dn, dk = _buf(dictionary)
    if out_cap is None:
        out_cap = max(64, n * 4)
        while True:
            r = inflate(data, window_bits, out_cap, dictionary)
            if r[0] == Z_BUF_ERROR and len(r[1]) == out_cap and out_cap < (1 << 31):
                out_cap *= 4
                continue
            return r
    out = (C.c_ubyte * max(out_cap, 1))()
    ol, used, ck = C.c_size_t(0), C.c_size_t(0), C.c_uint32(0)
    ret = lib().zo_inflate_oneshot(p, n, window_bits, dp, dn, C.addressof(out), out_cap, C.byref(ol),
                                   C.byref(used), C.byref(ck))
    return ret, bytes(out[: ol.value]) if ol.value else b"", used.value, ck.value


def deflate64_encode(data, max_len: int = 65538) -> bytes:
    """Raw deflate64 stream of `data` (test-side encoder, oracle/deflate64_enc.c)."""
    p, n, k = _buf(data)
    cap = n + n // 2 + 64
    out = (C.c_ubyte * cap)()
    L = lib()
    L.zo_deflate64_encode.restype = C.c_long
    L.zo_deflate64_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_int, C.c_void_p, C.c_size_t]
    r = L.zo_deflate64_encode(p, n, max_len, 1, C.addressof(out), cap)
    if r < 0:
        raise RuntimeError(f"deflate64 encoder failed: {r}")
    return bytes(out[:r])


def block_types(data, window_bits: int = -15):
    """Set of block types (0 stored, 1 fixed, 2 dynamic) of a complete stream."""
    ret, out, used, _ = inflate(data, window_bits)
    assert ret == Z_STREAM_END, ret
    counts = (C.c_uint32 * 3)()
    lib().zo_inflate_last_blocks(counts)
    return {t for t in range(3) if counts[t]}


def deflate(data, level: int = 6, wrap: int = 1, dictionary=None, flush: int = Z_FINISH, strategy: int = 0,
            rle_like_reference: bool = False) -> bytes:
    """One-shot deflateInit2_(level, 8, wbits(wrap), 8, strategy) [+ dictionary] + deflate(flush).
    Z_RLE is C zlib's algorithm unless rle_like_reference (the reference's scan never finds a run)."""
    p, n, k = _buf(data)
    dp, dn, dk = _buf(dictionary)
    cap = lib().zo_deflate_bound(n, wrap) + 64
    out = (C.c_ubyte * cap)()
    r = lib().zo_deflate_oneshot2(p, n, level, strategy, int(rle_like_reference), wrap, dp, dn, flush,
                                  C.addressof(out), cap)
    if r < 0:
        raise RuntimeError(f"oracle deflate failed: {r}")
    return bytes(out[:r])


def deflate_bound(n: int, wrap: int) -> int:
    return lib().zo_deflate_bound(n, wrap)


class InflateStream:
    """Thin handle over zo_inflate_stream for streaming-protocol tests."""

    def __init__(self, window_bits: int = 15):
        self._h = lib().zo_inflate_new()
        self.ret = lib().zo_inflate_init2(self._h, window_bits)
        self._keep = []

    def close(self):
        if self._h:
            lib().zo_inflate_free(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def set_dictionary(self, d: bytes) -> int:
        p, n, k = _buf(d)
        return lib().zo_inflate_set_dictionary(self._h, p, n)

    def reset(self) -> int:
        return lib().zo_inflate_reset(self._h)

    def step(self, data: bytes, out_cap: int, flush: int = Z_NO_FLUSH):
        """Feed `data`, allow `out_cap` output bytes.  Returns (ret, out, consumed)."""
        p, n, k = _buf(data)
        out = (C.c_ubyte * max(out_cap, 1))()
        L = lib()
        L.zo_inflate_set_input(self._h, p, n)
        L.zo_inflate_set_output(self._h, C.addressof(out), out_cap)
        ret = L.zo_inflate(self._h, flush)
        produced = out_cap - L.zo_inflate_avail_out(self._h)
        consumed = n - L.zo_inflate_avail_in(self._h)
        return ret, bytes(out[:produced]), consumed

    @property
    def total_in(self):
        return lib().zo_inflate_total_in(self._h)

    @property
    def total_out(self):
        return lib().zo_inflate_total_out(self._h)

    @property
    def adler(self):
        return lib().zo_inflate_adler(self._h)

    @property
    def msg(self):
        return (lib().zo_inflate_msg(self._h) or b"").decode()
