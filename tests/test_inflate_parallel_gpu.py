"""GPU parity of the segment-parallel decoder of ONE stream (csrc/zs_inflate_par.cu): whatever the stream --
its own STITCHED + SYNC deflate, C zlib streams with Z_SYNC_FLUSH / Z_FULL_FLUSH points, streams without any
flush point (cut at speculatively located block headers), corrupted, truncated, too little output room, stored data full of marker look-alikes -- the
result (output, total_in, status, message, checksum) equals the oracle restatement of inflate()
(src/mod/inflate/inflate.ts:332) and the engine's own serial decoder (ZS_INFLATE_SERIAL=1)."""
import os
import random
import zlib

import numpy as np
import pytest

from conftest import make_mixed, make_text, pkg, rand_bytes

pytestmark = pytest.mark.gpu


def _inflate(stream, wbits, cap, dictionary=None, serial=False):
    B = pkg("batch")
    if serial:
        os.environ["ZS_INFLATE_SERIAL"] = "1"
    try:
        return B.inflate_batch([stream], wbits, [cap], [dictionary] if dictionary is not None else None)
    finally:
        os.environ.pop("ZS_INFLATE_SERIAL", None)


def _check(oracle, stream, wbits, cap, dictionary=None, oracle_too=True):
    """parallel == serial (== oracle) on status, message, output, total_in, check."""
    p = _inflate(stream, wbits, cap, dictionary)
    s = _inflate(stream, wbits, cap, dictionary, serial=True)
    assert int(p.status[0]) == int(s.status[0]), (int(p.status[0]), int(s.status[0]), p.message(0), s.message(0))
    assert p.message(0) == s.message(0)
    assert p.output(0) == s.output(0), (len(p.output(0)), len(s.output(0)))
    assert int(p.in_used[0]) == int(s.in_used[0])
    assert int(p.checks[0]) == int(s.checks[0])
    if oracle_too:
        ret, out, used, check = oracle.inflate(stream, wbits, cap, dictionary)
        assert int(p.status[0]) == ret and p.output(0) == out
        if ret == oracle.Z_STREAM_END:
            assert int(p.in_used[0]) == used
            if wbits > 0:
                assert int(p.checks[0]) == check
    return p


def _flushed(data, level, wbits, step, mode, zdict=None):
    """C zlib stream with a flush point every `step` input bytes."""
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, 0, zdict) if zdict else zlib.compressobj(level, zlib.DEFLATED, wbits)
    parts = []
    for i in range(0, len(data), step):
        parts.append(co.compress(data[i:i + step]))
        if i + step < len(data):
            parts.append(co.flush(mode))
    parts.append(co.flush())
    return b"".join(parts)


@pytest.mark.parametrize("wrap,wbits", [(0, -15), (1, 15), (2, 31)])
def test_own_stitched_sync_stream(gpu_ctx, oracle, wrap, wbits):
    B = pkg("batch")
    data = make_mixed(24 << 20, 21)
    for level, chunk in ((6, 65536), (1, 262144)):
        r = B.deflate_batch(data, chunk, level, wrap, B.MODE_STITCHED, B.FLAG_SYNC)
        p = _check(oracle, r.data, wbits, len(data) + 64, oracle_too=(level == 6))
        assert int(p.status[0]) == 1 and p.output(0) == data and int(p.in_used[0]) == len(r.data)
        if wrap == 1:
            assert int(p.checks[0]) == zlib.adler32(data)
        if wrap == 2:
            assert int(p.checks[0]) == zlib.crc32(data)


def test_parallel_path_is_taken(gpu_ctx):
    """The planned segments are counted by the engine's profile: the decode kernel of the parallel path ran and
    the serial decoder had (almost) nothing left."""
    B = pkg("batch")
    data = make_text(16 << 20, 5)
    r = B.deflate_batch(data, 65536, 6, 1, B.MODE_STITCHED, B.FLAG_SYNC)
    gpu_ctx.profile(True)
    gpu_ctx.profile_read()
    p = _inflate(r.data, 15, len(data) + 64)
    prof = gpu_ctx.profile_read()
    gpu_ctx.profile(False)
    assert p.output(0) == data
    assert "par_decode_kernel" in prof and "par_window_kernel" in prof
    # 16 MiB through one warp takes > 1 s; the parallel decode of its ~256 segments a few ms
    assert prof["par_decode_kernel"][1] < 200.0 and prof["inflate_kernel"][1] < 100.0, prof


@pytest.mark.parametrize("mode", [zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])
def test_c_zlib_streams_with_flush_points(gpu_ctx, oracle, mode):
    data = make_text(6 << 20, 8) + make_mixed(6 << 20, 9)
    for wbits in (15, 31, -15):
        for level, step in ((6, 100000), (9, 37111), (1, 1 << 20)):
            z = _flushed(data, level, wbits, step, mode)
            p = _check(oracle, z, wbits, len(data) + 64)
            assert int(p.status[0]) == 1 and p.output(0) == data


def test_stream_without_flush_points_and_tiny_segments(gpu_ctx, oracle):
    data = make_text(3 << 20, 12)
    z = zlib.compress(data, 6)
    p = _check(oracle, z, 15, len(data) + 64)
    assert p.output(0) == data
    # a flush point every 100 bytes: thousands of candidates per tile, one kept per tile
    z = _flushed(data[: 1 << 20], 6, 15, 100, zlib.Z_SYNC_FLUSH)
    p = _check(oracle, z, 15, (1 << 20) + 64)
    assert p.output(0) == data[: 1 << 20]


def test_streams_without_flush_points_are_cut_speculatively(gpu_ctx, oracle):
    """Ordinary one-shot streams (zlib.compress / gzip / raw, every level incl. stored and fixed blocks): no
    marker to cut at, so the cuts are dynamic block headers found by par_spec_kernel.  Same results as the serial
    decoder and the oracle -- also when the stream is truncated or corrupted behind the first cuts, when the output
    room runs out, and with a preset dictionary -- and the parallel path really ran."""
    data = make_text(5 << 20, 31) + make_mixed(5 << 20, 32) + rand_bytes(300000, 33) + make_text(2 << 20, 34)
    for wbits in (15, 31, -15):
        for level, strategy in ((6, 0), (1, 0), (9, 0), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (0, 0)):
            co = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
            z = co.compress(data) + co.flush()
            p = _check(oracle, z, wbits, len(data) + 64, oracle_too=(level == 6 and strategy == 0))
            assert int(p.status[0]) == 1 and p.output(0) == data, (wbits, level, strategy)
    z = zlib.compress(data, 6)
    # the parallel kernels decoded it, the serial decoder only finished it
    gpu_ctx.profile(True)
    gpu_ctx.profile_read()
    p = _inflate(z, 15, len(data) + 64)
    prof = gpu_ctx.profile_read()
    gpu_ctx.profile(False)
    assert p.output(0) == data
    assert "par_spec_kernel" in prof and prof["par_decode_kernel"][1] < 100.0 and prof["inflate_kernel"][1] < 50.0, prof
    # damage behind the first cuts, truncation, too little room: verdict, message and output prefix of the serial decoder
    rng = random.Random(6)
    for _ in range(6):
        bad = bytearray(z)
        at = rng.randrange(len(z) // 3, len(z) - 8)
        bad[at] ^= 1 << rng.randrange(8)
        _check(oracle, bytes(bad), 15, len(data) + 64)
    for cut in (len(z) // 2, len(z) - 3, len(z) - 1):
        _check(oracle, z[:cut], 15, len(data) + 64)
    _check(oracle, z, 15, len(data) // 2)
    _check(oracle, z + b"trailing bytes", 15, len(data) + 64)
    # raw stream with a preset dictionary: the first segments reach into it
    zd = data[: 30000]
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, 0, zd)
    zr = co.compress(data[30000:]) + co.flush()
    p = _check(oracle, zr, -15, len(data), dictionary=zd)
    assert p.output(0) == data[30000:]


def test_marker_lookalikes_in_stored_data(gpu_ctx, oracle):
    """Stored blocks full of 00 00 FF FF: every tile gets a false candidate; the chain must not follow them."""
    rng = random.Random(4)
    blocks = []
    for _ in range(3000):
        k = rng.randrange(4)
        blocks.append(b"\x00\x00\xff\xff" * rng.randrange(1, 300) if k == 0 else rand_bytes(rng.randrange(1, 3000), rng.randrange(1 << 30))
                      if k == 1 else b"\x00\x00\xff\xff\x00\x00\x00\xff\xff" * 50 if k == 2 else bytes(rng.randrange(1, 2000)))
    data = b"".join(blocks)
    for level in (0, 1, 6):
        z = _flushed(data, level, 15, 50000, zlib.Z_SYNC_FLUSH)
        p = _check(oracle, z, 15, len(data) + 64)
        assert int(p.status[0]) == 1 and p.output(0) == data


def test_preset_dictionary_and_history(gpu_ctx, oracle):
    B = pkg("batch")
    data = make_text(9 << 20, 31)
    hist = data[: 32768]
    # a raw stream whose first segment matches into the preset dictionary
    z = _flushed(data[32768:], 6, -15, 200000, zlib.Z_SYNC_FLUSH, zdict=hist)
    p = _check(oracle, z, -15, len(data), dictionary=hist)
    assert int(p.status[0]) == 1 and p.output(0) == data[32768:]
    # the same without the dictionary: "invalid distance too far back", wherever the first such match is
    p = _check(oracle, z, -15, len(data))
    assert int(p.status[0]) == -3 and p.message(0) == "invalid distance too far back"
    # a short dictionary: only the distances beyond it fail
    p = _check(oracle, z, -15, len(data), dictionary=hist[-1000:])
    assert int(p.status[0]) == -3


def test_errors_truncation_and_short_output(gpu_ctx, oracle):
    B = pkg("batch")
    data = make_mixed(8 << 20, 3)
    good = B.deflate_batch(data, 65536, 6, 1, B.MODE_STITCHED, B.FLAG_SYNC).data
    rng = random.Random(9)
    # corrupted bytes anywhere
    for _ in range(12):
        bad = bytearray(good)
        at = rng.randrange(2, len(bad) - 4)
        bad[at] ^= 1 << rng.randrange(8)
        _check(oracle, bytes(bad), 15, len(data) + 64)
    # a wrong trailer
    bad = bytearray(good); bad[-1] ^= 0x55
    p = _check(oracle, bytes(bad), 15, len(data) + 64)
    assert int(p.status[0]) == -3 and p.message(0) == "incorrect data check"
    # truncated at every kind of place: mid segment, exactly behind a marker, inside the trailer
    cuts = [len(good) // 3, len(good) - 2, len(good) - 5]
    m = good.find(b"\x00\x00\xff\xff", len(good) // 2)
    cuts += [m + 4, m + 2, m + 5]
    for cut in cuts:
        p = _check(oracle, good[:cut], 15, len(data) + 64)
        assert int(p.status[0]) == -5
    # output room: too small by a lot, by one byte, exact
    for cap in (len(data) // 2, len(data) - 1, len(data)):
        p = _check(oracle, good, 15, cap)
        assert (int(p.status[0]) == 1) == (cap == len(data))
    # garbage behind the stream is not consumed
    p = _check(oracle, good + b"tail bytes", 15, len(data) + 64)
    assert int(p.status[0]) == 1 and int(p.in_used[0]) == len(good)


def test_stream_dev_entry_point(gpu_ctx):
    """zs_inflate_stream_dev: device buffers in, results on the device."""
    import ctypes as C
    import torch
    B, capi = pkg("batch"), pkg("capi")
    lib = capi.load()
    data = make_text(12 << 20, 77)
    r = B.deflate_batch(data, 65536, 6, 2, B.MODE_STITCHED, B.FLAG_SYNC)
    pad = bytes((-len(r.data)) % 8)
    d_in = torch.frombuffer(bytearray(r.data + pad), dtype=torch.uint8).cuda()
    d_out = torch.zeros(len(data) + 64, dtype=torch.uint8, device="cuda")
    res = torch.zeros(4, dtype=torch.int64, device="cuda")     # out_len, in_used, check, status
    rc = lib.zs_inflate_stream_dev(gpu_ctx.handle, d_in.data_ptr(), len(r.data), 31, d_out.data_ptr(), d_out.numel(),
                                   res.data_ptr(), res.data_ptr() + 8, res.data_ptr() + 16, res.data_ptr() + 24, None, 0)
    assert rc == 0, gpu_ctx.last_error()
    torch.cuda.synchronize()
    out_len, in_used, check, status = (int(x) for x in res.cpu())
    assert status & 0xffffffff == 1 and out_len == len(data) and in_used == len(r.data)
    assert check & 0xffffffff == zlib.crc32(data)
    assert bytes(d_out[:out_len].cpu().numpy()) == data


def test_decompression_stream_over_own_stream(gpu_ctx):
    """DecompressionStream (32 KiB slices) over a CompressionStream product: the shim batches its attempts, the
    attempts go through the parallel decoder."""
    S = pkg("streams")
    data = make_text(20 << 20, 44)
    comp = S.CompressionStream("gzip", {"level": 6}).transform(data)
    assert zlib.decompress(comp, 31) == data
    back = S.DecompressionStream("gzip").transform(comp)
    assert back == data
