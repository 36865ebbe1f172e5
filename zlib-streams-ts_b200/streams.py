"""CompressionStream / DecompressionStream over the GPU engine.

Mirror of src/mod/streams.ts: formats "deflate" | "gzip" | "deflate-raw" (| "deflate64-raw" for
decompression) map to windowBits 15 | 31 | -15 (| -16) exactly like streams.ts:220,233; input is
processed in slices of IN_CHUNK = 32 KiB (streams.ts:7) through output buffers of 64 KiB
(BufferPool, streams.ts:11-38); any return other than Z_OK / Z_STREAM_END raises
Error("process error: N") like streams.ts:116-118,169-181.  Python has no WHATWG TransformStream,
so the two classes expose the transformer protocol directly: write(chunk) -> list of output
chunks, close() -> list of output chunks (flush), and a one-shot transform(data).
"""
from __future__ import annotations

from . import zlib_api as Z

IN_CHUNK = 32 * 1024
OUT_CHUNK = 64 * 1024
_COMPRESS_WBITS = {"deflate": 15, "gzip": 31, "deflate-raw": -15}
_DECOMPRESS_WBITS = {"deflate": 15, "gzip": 31, "deflate-raw": -15, "deflate64-raw": -16}


class _ZlibTransform:
    """createZeroCopyZlibTransform({_createStream,_init,_process,_end}), streams.ts:40-45."""

    def __init__(self, create, init, process, end):
        self._create, self._init, self._process, self._end = create, init, process, end
        self._strm = None
        self._closed = False

    def _ensure(self):
        if self._strm is None:
            self._strm = self._create()
            rc = self._init(self._strm)
            if rc != Z.Z_OK:
                raise RuntimeError(f"init failed: {rc}")

    def write(self, chunk) -> list:
        if self._closed:
            raise RuntimeError("stream is closed")
        self._ensure()
        data = bytes(chunk)
        out = []
        strm = self._strm
        for lo in range(0, len(data), IN_CHUNK):
            piece = data[lo: lo + IN_CHUNK]
            strm.next_in, strm.next_in_index, strm.avail_in = piece, 0, len(piece)
            while strm.avail_in > 0:
                buf = bytearray(OUT_CHUNK)
                strm.next_out, strm.next_out_index, strm.avail_out = buf, 0, OUT_CHUNK
                rc = self._process(strm, Z.Z_NO_FLUSH)
                if rc not in (Z.Z_OK, Z.Z_STREAM_END):
                    raise RuntimeError(f"process error: {rc}")
                if strm.next_out_index:
                    out.append(bytes(buf[: strm.next_out_index]))
                if rc == Z.Z_STREAM_END:
                    break
        return out

    def close(self) -> list:
        if self._closed:
            return []
        self._ensure()
        self._closed = True
        out = []
        strm = self._strm
        strm.next_in, strm.next_in_index, strm.avail_in = b"", 0, 0
        while True:
            buf = bytearray(OUT_CHUNK)
            strm.next_out, strm.next_out_index, strm.avail_out = buf, 0, OUT_CHUNK
            rc = self._process(strm, Z.Z_FINISH)
            if strm.next_out_index:
                out.append(bytes(buf[: strm.next_out_index]))
            if rc == Z.Z_STREAM_END:
                break
            if rc != Z.Z_OK:
                raise RuntimeError(f"finalization error: {rc}")
        rc = self._end(strm)
        if rc != Z.Z_OK:
            raise RuntimeError(f"end failed: {rc}")
        return out

    def transform(self, data) -> bytes:
        return b"".join(self.write(data) + self.close())


class CompressionStream(_ZlibTransform):
    """new CompressionStream(format, {level}), streams.ts:242-251."""

    def __init__(self, format: str, options: dict | None = None):
        if format not in _COMPRESS_WBITS:
            raise TypeError(f"Unsupported compression format: {format}")
        level = (options or {}).get("level", -1)
        wbits = _COMPRESS_WBITS[format]
        super().__init__(Z.createDeflateStream, lambda s: Z.deflateInit2_(s, level, Z.Z_DEFLATED, wbits, 8, 0),
                         Z.deflate, Z.deflateEnd)


class DecompressionStream(_ZlibTransform):
    """new DecompressionStream(format), streams.ts:253-262."""

    def __init__(self, format: str):
        if format not in _DECOMPRESS_WBITS:
            raise TypeError(f"Unsupported compression format: {format}")
        wbits = _DECOMPRESS_WBITS[format]
        super().__init__(Z.createInflateStream, lambda s: Z.inflateInit2_(s, wbits), Z.inflate, Z.inflateEnd)


CompressionStreamZlib = CompressionStream
DecompressionStreamZlib = DecompressionStream
