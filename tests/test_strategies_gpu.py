"""GPU: level 0 and the non-default strategies (scope row f4).  The reference pins round trips
(test/deflate/test-deflate-stored.ts, test-deflate-huff.ts, test-deflate-rle.ts,
test-deflate-rle-huff-edge.ts, test-deflate-flush-strategy-branches.ts, coverage-deflate-params.spec.ts);
sizes are compared with C zlib run with the same level / strategy (the reference's own cross-oracle)."""
import zlib

import pytest

from conftest import make_mixed, make_text, pkg, rand_bytes

pytestmark = pytest.mark.gpu

WB = {0: -15, 1: 15, 2: 31}
STRATS = {"filtered": 1, "huffman_only": 2, "rle": 3, "fixed": 4}


def _zlib_size(data, level, strategy, wbits=-15):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
    return len(co.compress(data) + co.flush())


def _rle_data(n):
    # the reference's RLE test pattern (test-deflate-rle.ts:11-14)
    return bytes(0xaa if i % 128 < 120 else i & 0xff for i in range(n))


DATA = {
    "text": lambda: make_text(400000, 31),
    "mixed": lambda: make_mixed(600000, 32),
    "rle": lambda: _rle_data(300000),
    "bytes": lambda: bytes(i & 0xff for i in range(70000)),   # test-deflate-huff.ts:10-14
    "empty": lambda: b"",
}


@pytest.mark.parametrize("name", list(DATA))
def test_level0_stored(gpu_ctx, oracle, name):
    B = pkg("batch")
    data = DATA[name]()
    for wrap in (0, 1, 2):
        for mode in (B.MODE_INDEPENDENT, B.MODE_STITCHED):
            r = B.deflate_batch(data, 100000, 0, wrap, mode)
            if mode == B.MODE_STITCHED:
                s = r.data
                ret, out, used, _ = oracle.inflate(s, WB[wrap], len(data) + 64)
                assert ret == oracle.Z_STREAM_END and out == data and used == len(s)
                d = zlib.decompressobj(WB[wrap])
                assert d.decompress(s) + d.flush() == data and d.eof
                # stored blocks: 5 bytes per block of <= 65535 bytes, like deflate_stored
                ref = len(zlib.compressobj(0, zlib.DEFLATED, WB[wrap]).compress(data)) if data else 0
                assert len(s) <= max(ref, len(data)) + 5 * (len(data) // 65535 + 2 + len(data) // 100000) + 24
            else:
                for i in range(max(1, -(-len(data) // 100000))):
                    chunk = data[i * 100000: (i + 1) * 100000]
                    d = zlib.decompressobj(WB[wrap])
                    assert d.decompress(r.stream(i)) + d.flush() == chunk and d.eof


@pytest.mark.parametrize("strategy", list(STRATS))
@pytest.mark.parametrize("level", [1, 6])
@pytest.mark.parametrize("name", ["text", "mixed", "rle", "bytes", "empty"])
def test_strategy_roundtrip_and_size(gpu_ctx, oracle, name, level, strategy):
    B = pkg("batch")
    data = DATA[name]()
    st = STRATS[strategy]
    r = B.deflate_batch(data, 65536, level, 1, B.MODE_STITCHED, flags=B.flag_strategy(st))
    s = r.data
    ret, out, used, check = oracle.inflate(s, 15, len(data) + 64)
    assert ret == oracle.Z_STREAM_END and out == data and used == len(s)
    assert zlib.decompress(s) == data
    ref = _zlib_size(data, level, st, 15)
    assert len(oracle.deflate(data, level, 1, strategy=st)) == ref   # the oracle's restatement is C zlib's, byte for byte
    # every 64 KiB chunk starts its own block: allow one dynamic header per chunk on top of the 3 %
    n_chunks = max(1, -(-len(data) // 65536))
    assert len(s) <= ref * 1.03 + 64 + 80 * n_chunks, (name, level, strategy, len(s), ref)
    if strategy == "huffman_only" and len(data) > 1000:
        # literals only: never smaller than the entropy-coded bytes of zlib's Z_HUFFMAN_ONLY by much
        assert len(s) >= ref * 0.97
    if strategy == "fixed" and data:
        # no dynamic block anywhere: walk the block headers with the oracle's block lister
        types = oracle.block_types(s, 15)
        assert 2 not in types


def test_streaming_api_strategies(gpu_ctx):
    """test-deflate-huff.ts / test-deflate-rle.ts / test-deflate-stored.ts through the z_stream API."""
    Z = pkg("zlib_api")
    from test_zlib_api_gpu import chunked_inflate
    cases = [(bytes(i & 0xff for i in range(16 * 1024)), 6, Z.Z_HUFFMAN_ONLY), (_rle_data(64 * 1024), 6, Z.Z_RLE),
             (make_text(50000, 5), 0, Z.Z_DEFAULT_STRATEGY), (make_text(50000, 6), 6, Z.Z_FIXED), (make_text(50000, 7), 6, Z.Z_FILTERED)]
    for data, level, strategy in cases:
        s = Z.createDeflateStream()
        assert Z.deflateInit2_(s, level, Z.Z_DEFLATED, 15, 8, strategy) == Z.Z_OK
        out = bytearray()
        for pos in range(0, len(data), 4096):
            piece = data[pos: pos + 4096]
            s.next_in, s.next_in_index, s.avail_in = piece, 0, len(piece)
            while s.avail_in:
                buf = bytearray(1024)
                s.next_out, s.next_out_index, s.avail_out = buf, 0, 1024
                assert Z.deflate(s, Z.Z_NO_FLUSH) == Z.Z_OK
                out += buf[: s.next_out_index]
        while True:
            buf = bytearray(1024)
            s.next_in, s.next_in_index, s.avail_in = b"", 0, 0
            s.next_out, s.next_out_index, s.avail_out = buf, 0, 1024
            r = Z.deflate(s, Z.Z_FINISH)
            out += buf[: s.next_out_index]
            if r == Z.Z_STREAM_END:
                break
            assert r == Z.Z_OK
        assert Z.deflateEnd(s) == Z.Z_OK
        assert zlib.decompress(bytes(out)) == data
        got, r, _ = chunked_inflate(bytes(out), 15, 8, 16) if len(data) <= 20000 else chunked_inflate(bytes(out), 15, 4096, 8192)
        assert r == Z.Z_STREAM_END and got == data


def test_params_pending_reset(gpu_ctx):
    """deflateParams mid-stream, deflatePending / deflateUsed on a fresh stream
    (test-deflatePending-used.ts:16-30), deflateReset reuse."""
    Z = pkg("zlib_api")
    s = Z.createDeflateStream()
    assert Z.deflateInit(s, 6) == Z.Z_OK
    p, b = Z.Ref(7), Z.Ref(7)
    assert Z.deflatePending(s, p, b) == Z.Z_OK and p._value == 0 and b._value == 0
    assert Z.deflateUsed(s, b) == Z.Z_OK and b._value == 0
    assert Z.deflatePending(None) == Z.Z_STREAM_ERROR and Z.deflateParams(None, 1, 0) == Z.Z_STREAM_ERROR
    a, c, e = make_text(30000, 1), rand_bytes(20000, 2), make_text(30000, 3)
    out = bytearray()

    def feed(piece, flush):
        s.next_in, s.next_in_index, s.avail_in = piece, 0, len(piece)
        while True:
            buf = bytearray(65536)
            s.next_out, s.next_out_index, s.avail_out = buf, 0, len(buf)
            r = Z.deflate(s, flush)
            out.extend(buf[: s.next_out_index])
            if (flush == Z.Z_FINISH and r == Z.Z_STREAM_END) or (flush != Z.Z_FINISH and s.avail_in == 0 and s.avail_out):
                return r
            assert r == Z.Z_OK

    def params(level, strategy):
        while True:
            buf = bytearray(65536)
            s.next_in, s.next_in_index, s.avail_in = b"", 0, 0
            s.next_out, s.next_out_index, s.avail_out = buf, 0, len(buf)
            r = Z.deflateParams(s, level, strategy)
            out.extend(buf[: s.next_out_index])
            if r != Z.Z_BUF_ERROR:
                return r

    feed(a, Z.Z_NO_FLUSH)
    assert params(0, Z.Z_DEFAULT_STRATEGY) == Z.Z_OK        # compressible part done at level 6, now store
    feed(c, Z.Z_NO_FLUSH)
    assert params(9, Z.Z_HUFFMAN_ONLY) == Z.Z_OK
    assert Z.deflateParams(s, 10, 0) == Z.Z_STREAM_ERROR and Z.deflateParams(s, 1, 5) == Z.Z_STREAM_ERROR
    assert feed(e, Z.Z_FINISH) == Z.Z_STREAM_END
    assert zlib.decompress(bytes(out)) == a + c + e
    # reuse after deflateReset: same parameters, fresh stream
    assert Z.deflateReset(s) == Z.Z_OK and s.total_in == 0 and s.total_out == 0
    out.clear()
    assert feed(a, Z.Z_FINISH) == Z.Z_STREAM_END
    assert zlib.decompress(bytes(out)) == a
    assert Z.deflateEnd(s) == Z.Z_OK


def test_gzip_header_fields(gpu_ctx):
    """deflateSetHeader / inflateGetHeader (scope row f3; test-deflate-gzip-header.ts, test/inflate/test-header.ts):
    the header written for the caller's fields is what C zlib (gzip module semantics) and our own
    inflateGetHeader read back."""
    import gzip, io
    Z = pkg("zlib_api")
    data = make_text(40000, 9)
    for hcrc in (0, 1):
        s = Z.createDeflateStream()
        assert Z.deflateInit2_(s, 6, Z.Z_DEFLATED, 31, 8, 0) == Z.Z_OK
        head = Z.GzipHeader(text=1, time=0x5f3759df, os=3, extra=b"\x01\x02abcd", name=b"file.txt", comment=b"a comment", hcrc=hcrc)
        assert Z.deflateSetHeader(s, head) == Z.Z_OK
        out = bytearray()
        s.next_in, s.next_in_index, s.avail_in = data, 0, len(data)
        while True:
            buf = bytearray(4096)
            s.next_out, s.next_out_index, s.avail_out = buf, 0, len(buf)
            r = Z.deflate(s, Z.Z_FINISH)
            out += buf[: s.next_out_index]
            if r == Z.Z_STREAM_END:
                break
            assert r == Z.Z_OK
        assert Z.deflateEnd(s) == Z.Z_OK
        out = bytes(out)
        assert out[:4] == bytes([0x1f, 0x8b, 8, 1 + 2 * hcrc + 4 + 8 + 16]) and out[9] == 3
        assert int.from_bytes(out[4:8], "little") == 0x5f3759df
        g = gzip.GzipFile(fileobj=io.BytesIO(out))
        assert g.read() == data and g.mtime == 0x5f3759df
        assert zlib.decompress(out, 31) == data
        # and back through inflateGetHeader, fed in small pieces
        t = Z.createInflateStream()
        assert Z.inflateInit2_(t, 31) == Z.Z_OK
        got_head = Z.GzipHeader(extra_max=64, name_max=64, comm_max=64)
        assert Z.inflateGetHeader(t, got_head) == Z.Z_OK and got_head._done == 0
        res = bytearray()
        pos, r = 0, Z.Z_OK
        while r != Z.Z_STREAM_END:
            piece = out[pos: pos + 7]
            t.next_in, t.next_in_index, t.avail_in = piece, 0, len(piece)
            buf = bytearray(65536)
            t.next_out, t.next_out_index, t.avail_out = buf, 0, len(buf)
            r = Z.inflate(t, Z.Z_NO_FLUSH)
            assert r in (Z.Z_OK, Z.Z_STREAM_END, Z.Z_BUF_ERROR), (r, t.msg)
            res += buf[: t.next_out_index]
            pos += len(piece) - t.avail_in
        assert bytes(res) == data
        assert got_head._done == 1 and got_head._text == 1 and got_head._time == 0x5f3759df and got_head._os == 3
        assert got_head._extra == b"\x01\x02abcd" and got_head._extra_len == 6
        assert got_head._name == b"file.txt" and got_head._comment == b"a comment" and got_head._hcrc == hcrc
        assert Z.inflateEnd(t) == Z.Z_OK
    # not a gzip stream: Z_STREAM_ERROR for zlib-only windowBits, done = -1 under auto-detection
    t = Z.createInflateStream()
    assert Z.inflateInit2_(t, 15) == Z.Z_OK
    assert Z.inflateGetHeader(t, Z.GzipHeader()) == Z.Z_STREAM_ERROR
    assert Z.inflateReset2(t, 47) == Z.Z_OK
    h = Z.GzipHeader()
    assert Z.inflateGetHeader(t, h) == Z.Z_OK
    z = zlib.compress(data)
    t.next_in, t.next_in_index, t.avail_in = z, 0, len(z)
    buf = bytearray(len(data) + 64)
    t.next_out, t.next_out_index, t.avail_out = buf, 0, len(buf)
    assert Z.inflate(t, Z.Z_FINISH) == Z.Z_STREAM_END and bytes(buf[: t.next_out_index]) == data
    assert h._done == -1
    assert Z.inflateEnd(t) == Z.Z_OK
    # deflateSetHeader needs a gzip stream
    s = Z.createDeflateStream()
    assert Z.deflateInit(s, 6) == Z.Z_OK
    assert Z.deflateSetHeader(s, Z.GzipHeader()) == Z.Z_STREAM_ERROR
    assert Z.deflateEnd(s) == Z.Z_OK
