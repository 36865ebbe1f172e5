"""Host-side mirror of the reference's zlib-style API over the GPU engine.

Same names, argument meaning, return codes and draining protocol as
src/mod/deflate/deflate.ts (createDeflateStream :80, deflateInit :238, deflateInit2_ :253, deflate
:716, deflateEnd :991, deflateSetDictionary :367, deflateBound :615) and src/mod/inflate/inflate.ts
(createInflateStream :68, inflateInit :74, inflateInit2_ :174, inflate :332, inflateEnd :1187,
inflateSetDictionary :1220, inflateReset :124).  The data carrier is the reference's Stream
(src/mod/common/types.ts:1-15): next_in / next_in_index / avail_in / total_in, next_out /
next_out_index / avail_out / total_out, msg, _adler, _data_type, _state.

All work is delegated to the C ABI streaming shim (zs_stream_* in include/zsgpu.h); this module is
marshalling only -- it is what the TypeScript facade of INTEGRATION.md does through the addon.
"""
from __future__ import annotations

import ctypes as C

from . import capi
from .capi import (Z_BLOCK, Z_BUF_ERROR, Z_DATA_ERROR, Z_FINISH, Z_FULL_FLUSH, Z_MEM_ERROR, Z_NEED_DICT,  # noqa: F401
                   Z_NO_FLUSH, Z_OK, Z_PARTIAL_FLUSH, Z_STREAM_END, Z_STREAM_ERROR, Z_SYNC_FLUSH)

Z_DEFLATED = 8
Z_DEFAULT_COMPRESSION = -1
Z_DEFAULT_STRATEGY = 0
Z_FILTERED = 1
Z_HUFFMAN_ONLY = 2
Z_RLE = 3
Z_FIXED = 4
Z_NO_COMPRESSION = 0
Z_BEST_SPEED = 1
Z_BEST_COMPRESSION = 9
DEF_WBITS = 15
DEF_MEM_LEVEL = 8


class Stream:
    """The reference's Stream object (createStream, src/mod/common/utils.ts:44-60)."""

    def __init__(self):
        self.next_in = b""
        self.next_in_index = 0
        self.avail_in = 0
        self.total_in = 0
        self.next_out = bytearray(0)
        self.next_out_index = 0
        self.avail_out = 0
        self.total_out = 0
        self.msg = ""
        self._data_type = 0
        self._adler = 0
        self._state = None      # ("deflate" | "inflate", ZStream)
        self._gzhead = None     # inflateGetHeader request
        self._keep = None


def _ctx():
    from .batch import default_context
    return default_context()


def createDeflateStream() -> Stream:
    return Stream()


def createInflateStream() -> Stream:
    return Stream()


def _kind(strm, kind):
    return strm is not None and isinstance(strm, Stream) and strm._state is not None and strm._state[0] == kind


def deflateInit(strm: Stream, level: int) -> int:
    return deflateInit2_(strm, level)


def deflateInit2_(strm: Stream, level: int, method: int = Z_DEFLATED, windowBits: int = DEF_WBITS,
                  memLevel: int = DEF_MEM_LEVEL, strategy: int = Z_DEFAULT_STRATEGY) -> int:
    if strm is None:
        return Z_STREAM_ERROR
    strm.msg = ""
    zs = capi.ZStream()
    rc = capi.load().zs_stream_deflate_init(_ctx().handle, C.byref(zs), level, method, windowBits, memLevel, strategy)
    if rc != Z_OK:
        strm.msg = (zs.msg or b"").decode() if zs.msg else ""
        return rc
    strm._state = ("deflate", zs)
    strm.total_in = strm.total_out = 0
    strm._adler = zs.adler
    strm._data_type = zs.data_type
    return Z_OK


deflateInit2 = deflateInit2_


def deflateSetDictionary(strm: Stream, dictionary, dictLength: int | None = None) -> int:
    if not _kind(strm, "deflate") or dictionary is None:
        return Z_STREAM_ERROR
    d = bytes(dictionary[: dictLength] if dictLength is not None else dictionary)
    buf = (C.c_ubyte * max(len(d), 1)).from_buffer_copy(d or b"\0")
    zs = strm._state[1]
    rc = capi.load().zs_stream_deflate_set_dictionary(C.byref(zs), buf, len(d))
    strm._adler = zs.adler
    return rc


def _call(strm: Stream, fn, flush: int) -> int:
    zs = strm._state[1]
    n_in, n_out = strm.avail_in, strm.avail_out
    src = bytes(strm.next_in[strm.next_in_index: strm.next_in_index + n_in]) if n_in else b""
    in_buf = (C.c_ubyte * max(n_in, 1)).from_buffer_copy(src or b"\0")
    out_buf = (C.c_ubyte * max(n_out, 1))()
    zs.next_in = C.addressof(in_buf)
    zs.avail_in = n_in
    zs.next_out = C.addressof(out_buf)
    zs.avail_out = n_out
    rc = fn(C.byref(zs), flush)
    used = n_in - zs.avail_in
    made = n_out - zs.avail_out
    if made:
        strm.next_out[strm.next_out_index: strm.next_out_index + made] = bytes(out_buf[:made])
    strm.next_in_index += used
    strm.avail_in -= used
    strm.next_out_index += made
    strm.avail_out -= made
    strm.total_in = zs.total_in
    strm.total_out = zs.total_out
    strm._adler = zs.adler
    strm._data_type = zs.data_type
    strm.msg = (zs.msg or b"").decode() if zs.msg else ""
    return rc


def deflate(strm: Stream, flush: int) -> int:
    if not _kind(strm, "deflate"):
        return Z_STREAM_ERROR
    if strm.next_out is None or (strm.avail_in != 0 and strm.next_in is None):
        return Z_STREAM_ERROR
    return _call(strm, capi.load().zs_stream_deflate, flush)


def deflateEnd(strm: Stream) -> int:
    if not _kind(strm, "deflate"):
        return Z_STREAM_ERROR
    rc = capi.load().zs_stream_deflate_end(C.byref(strm._state[1]))
    strm._state = None
    return rc


class GzipHeader:
    """GzipHeader of src/mod/common/types.ts:137-151."""

    def __init__(self, text=0, time=0, xflags=0, os=0, extra=None, extra_len=0, name=None, comment=None, hcrc=0,
                 extra_max=0, name_max=0, comm_max=0):
        self._text, self._time, self._xflags, self._os = text, time, xflags, os
        self._extra, self._extra_len, self._extra_max = extra, extra_len if extra_len else (len(extra) if extra else 0), extra_max
        self._name, self._name_max = name, name_max
        self._comment, self._comm_max = comment, comm_max
        self._hcrc = hcrc
        self._done = 0


def deflateSetHeader(strm: Stream, head: GzipHeader) -> int:
    """deflate.ts:497-503"""
    if not _kind(strm, "deflate") or head is None:
        return Z_STREAM_ERROR
    g = capi.GzHeader()
    keep = []

    def buf(b, terminate):
        if b is None:
            return None
        raw = bytes(b)
        if terminate and (not raw or raw[-1] != 0):
            raw = raw.split(b"\0")[0] + b"\0"
        a = (C.c_ubyte * max(len(raw), 1)).from_buffer_copy(raw or b"\0")
        keep.append(a)
        return C.addressof(a)

    g.text, g.time, g.xflags, g.os, g.hcrc = int(bool(head._text)), head._time & 0xffffffff, head._xflags, head._os & 0xff, int(bool(head._hcrc))
    g.extra = buf(head._extra, False)
    g.extra_len = head._extra_len
    g.name = buf(head._name, True)
    g.comment = buf(head._comment, True)
    return capi.load().zs_stream_deflate_set_header(C.byref(strm._state[1]), C.byref(g))


def inflateGetHeader(strm: Stream, head: GzipHeader) -> int:
    """inflate.ts (inflateGetHeader): `head` is filled while inflate() walks the gzip header."""
    if not _kind(strm, "inflate") or head is None:
        return Z_STREAM_ERROR
    g = capi.GzHeader()
    bufs = {}
    for field, cap in (("extra", head._extra_max), ("name", head._name_max), ("comment", head._comm_max)):
        if cap:
            bufs[field] = (C.c_ubyte * cap)()
            setattr(g, field, C.addressof(bufs[field]))
    g.extra_max, g.name_max, g.comm_max = head._extra_max, head._name_max, head._comm_max
    rc = capi.load().zs_stream_inflate_get_header(C.byref(strm._state[1]), C.byref(g))
    if rc == Z_OK:
        head._done = 0
        strm._gzhead = (head, g, bufs)
    return rc


def _sync_gzhead(strm: Stream):
    gz = getattr(strm, "_gzhead", None)
    if not gz:
        return
    head, g, bufs = gz
    if g.done == 0 or head._done != 0:
        return
    head._done = g.done
    if g.done == 1:
        head._text, head._time, head._xflags, head._os, head._hcrc = g.text, g.time, g.xflags, g.os, g.hcrc
        head._extra_len = g.extra_len
        if "extra" in bufs:
            head._extra = bytes(bufs["extra"][: min(g.extra_len, head._extra_max)])
        if "name" in bufs:
            head._name = bytes(bufs["name"]).split(b"\0")[0]
        if "comment" in bufs:
            head._comment = bytes(bufs["comment"]).split(b"\0")[0]


class Ref:
    """The reference's out-parameter box ({ _value: number })."""

    def __init__(self, value: int = 0):
        self._value = value


def deflateResetKeep(strm: Stream) -> int:
    """deflate.ts:444-487"""
    if not _kind(strm, "deflate"):
        return Z_STREAM_ERROR
    zs = strm._state[1]
    rc = capi.load().zs_stream_deflate_reset(C.byref(zs))
    strm.total_in = strm.total_out = 0
    strm.msg = ""
    strm._adler = zs.adler
    strm._data_type = zs.data_type
    return rc


deflateReset = deflateResetKeep   # deflate.ts:489-495: ResetKeep + lm_init; the engine keeps no match state


def deflateParams(strm: Stream, level: int, strategy: int) -> int:
    """deflate.ts:553-595.  Pending input is compressed with the old parameters first; Z_BUF_ERROR
    asks the caller to drain the output and call again, as in the reference."""
    if not _kind(strm, "deflate"):
        return Z_STREAM_ERROR
    lib = capi.load()
    return _call(strm, lambda zs, _f: lib.zs_stream_deflate_params(zs, level, strategy), 0)


def deflatePending(strm: Stream, pending: Ref | None = None, bits: Ref | None = None) -> int:
    """deflate.ts:505-516"""
    if not _kind(strm, "deflate"):
        return Z_STREAM_ERROR
    p, b = C.c_uint32(0), C.c_int(0)
    rc = capi.load().zs_stream_deflate_pending(C.byref(strm._state[1]), C.byref(p), C.byref(b))
    if pending is not None:
        pending._value = p.value
    if bits is not None:
        bits._value = b.value
    return rc


def deflateUsed(strm: Stream, bits: Ref | None = None) -> int:
    """deflate.ts:518-526: bits used in the last byte written (every part ends byte aligned: 0)."""
    if not _kind(strm, "deflate"):
        return Z_STREAM_ERROR
    if bits is not None:
        bits._value = 0
    return Z_OK


def deflateBound(strm, sourceLen: int) -> int:
    """deflate.ts:615-674 for the default windowBits 15 / memLevel 8 state; like the reference it
    also answers for a null stream (conservative bound + 18)."""
    n = sourceLen
    if not _kind(strm, "deflate"):
        fixedlen = n + (n >> 3) + (n >> 8) + (n >> 9) + 4
        storelen = n + (n >> 5) + (n >> 7) + (n >> 11) + 7
        return max(fixedlen, storelen) + 18
    return n + (n >> 12) + (n >> 14) + (n >> 25) + 13 - 6 + 18


def inflateInit(strm: Stream) -> int:
    return inflateInit2_(strm, DEF_WBITS)


def inflateInit2_(strm: Stream, windowBits: int) -> int:
    if strm is None:
        return Z_STREAM_ERROR
    strm.msg = ""
    zs = capi.ZStream()
    rc = capi.load().zs_stream_inflate_init(_ctx().handle, C.byref(zs), windowBits)
    if rc != Z_OK:
        return rc
    strm._state = ("inflate", zs)
    strm.total_in = strm.total_out = 0
    strm._adler = zs.adler
    return Z_OK


inflateInit2 = inflateInit2_


def inflateSetDictionary(strm: Stream, dictionary, dictLength: int | None = None) -> int:
    if not _kind(strm, "inflate"):
        return Z_STREAM_ERROR
    d = bytes(dictionary[: dictLength] if dictLength is not None else dictionary)
    buf = (C.c_ubyte * max(len(d), 1)).from_buffer_copy(d or b"\0")
    return capi.load().zs_stream_inflate_set_dictionary(C.byref(strm._state[1]), buf, len(d))


def inflate(strm: Stream, flush: int) -> int:
    if not _kind(strm, "inflate"):
        return Z_STREAM_ERROR
    if strm.next_out is None or (strm.next_in is None and strm.avail_in != 0):
        return Z_STREAM_ERROR
    rc = _call(strm, capi.load().zs_stream_inflate, flush)
    _sync_gzhead(strm)
    return rc


def inflateReset(strm: Stream) -> int:
    if not _kind(strm, "inflate"):
        return Z_STREAM_ERROR
    zs = strm._state[1]
    rc = capi.load().zs_stream_inflate_reset(C.byref(zs))
    strm.total_in = strm.total_out = 0
    strm.msg = ""
    strm._gzhead = None
    return rc


inflateResetKeep = inflateReset   # inflate.ts:96-122: the engine has no window to keep


def inflateReset2(strm: Stream, windowBits: int) -> int:
    """inflate.ts:138-172"""
    if not _kind(strm, "inflate"):
        return Z_STREAM_ERROR
    zs = strm._state[1]
    rc = capi.load().zs_stream_inflate_reset2(C.byref(zs), windowBits)
    if rc == Z_OK:
        strm.total_in = strm.total_out = 0
        strm.msg = ""
        strm._adler = zs.adler
        strm._gzhead = None
    return rc


def inflateEnd(strm: Stream) -> int:
    if not _kind(strm, "inflate"):
        return Z_STREAM_ERROR
    rc = capi.load().zs_stream_inflate_end(C.byref(strm._state[1]))
    strm._state = None
    return rc
