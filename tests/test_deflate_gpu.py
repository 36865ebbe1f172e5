"""GPU parity: K1-K5 + K9 deflate.  The GPU bit stream is allowed to differ from the reference's
(chunk-parallel matching), so parity is: (1) every stream decodes bit-exactly to its input through
the oracle restatement of the reference's inflate AND through C zlib, (2) checksums / framing
bytes are bit-exact, (3) compressed size stays within 3 % of the reference (oracle deflate, which
is byte-exact with C zlib 1.3) at the same level."""
import zlib

import numpy as np
import pytest

from conftest import make_mixed, make_text, pkg, rand_bytes

pytestmark = pytest.mark.gpu

RATIO_TOL = 1.03
WB = {0: -15, 1: 15, 2: 31}


def _decode_both(oracle, stream, wrap, n, dictionary=None):
    ret, out, used, check = oracle.inflate(stream, WB[wrap], n + 64, dictionary)
    assert ret == oracle.Z_STREAM_END, (ret, len(out))
    assert used == len(stream)
    d = zlib.decompressobj(WB[wrap], zdict=dictionary) if dictionary else zlib.decompressobj(WB[wrap])
    assert d.decompress(stream) + d.flush() == out and d.eof
    return out


CORPORA = {
    "text": lambda: make_text(700000, 21),
    "mixed": lambda: make_mixed(900000, 22),
    "random": lambda: rand_bytes(300000, 23),
    "zeros": lambda: bytes(500000),
    "ramp": lambda: bytes(j % 251 for j in range(400000)),
    "tiny": lambda: b"hello, hello, hello!",
    "one": lambda: b"a",
    "empty": lambda: b"",
}


@pytest.mark.parametrize("name", list(CORPORA))
@pytest.mark.parametrize("level", [1, 6])
def test_independent_streams_roundtrip(gpu_ctx, oracle, name, level):
    B = pkg("batch")
    data = CORPORA[name]()
    for wrap in (0, 1, 2):
        r = B.deflate_batch(data, 65536, level, wrap, B.MODE_INDEPENDENT)
        n_chunks = max(1, -(-len(data) // 65536))
        assert r.out_off.size == n_chunks + 1 and int(r.out_off[-1]) == r.total_out_bytes == len(r.data)
        got = b""
        for i in range(n_chunks):
            chunk = data[i * 65536: (i + 1) * 65536]
            s = r.stream(i)
            assert _decode_both(oracle, s, wrap, len(chunk)) == chunk, (name, wrap, i)
            if wrap == 1:
                assert int(r.checks[i]) == zlib.adler32(chunk)
                assert s[:2] == {1: b"\x78\x01", 6: b"\x78\x9c"}[level]
            if wrap == 2:
                assert int(r.checks[i]) == zlib.crc32(chunk)
                assert s[:4] == b"\x1f\x8b\x08\x00" and s[9] == 255   # OS byte of the reference
            assert len(s) <= oracle.deflate_bound(len(chunk), wrap) + 8
            got += chunk
        assert got == data
        if wrap:
            want = zlib.adler32(data) if wrap == 1 else zlib.crc32(data)
            assert r.check == want


@pytest.mark.parametrize("level", [1, 3, 4, 6, 9])
def test_stitched_single_stream(gpu_ctx, oracle, level):
    B = pkg("batch")
    for name in ("text", "mixed", "random", "zeros", "empty", "one"):
        data = CORPORA[name]()
        for wrap, chunk in ((1, 65536), (2, 262144), (0, 16384)):
            r = B.deflate_batch(data, chunk, level, wrap, B.MODE_STITCHED)
            assert _decode_both(oracle, r.data, wrap, len(data)) == data, (name, wrap, level)
            # per-chunk bit lengths sum to the body length (the exclusive scan of the north star)
            hdr = {0: 0, 1: 16, 2: 80}[wrap]
            assert int(r.out_off[0]) == hdr
            assert (r.out_off[1:] - r.out_off[:-1] == r.out_bits).all()


def test_stitched_sync_markers(gpu_ctx, oracle):
    B = pkg("batch")
    data = make_text(300000, 5) + rand_bytes(100000, 6) + make_text(100000, 7)
    r = B.deflate_batch(data, 65536, 6, 1, B.MODE_STITCHED, B.FLAG_SYNC)
    assert _decode_both(oracle, r.data, 1, len(data)) == data
    # every chunk but the last ends byte aligned with 00 00 ff ff
    for i in range(1, r.out_off.size - 1):
        end = int(r.out_off[i])
        assert end % 8 == 0 and r.data[end // 8 - 4: end // 8] == b"\x00\x00\xff\xff"


@pytest.mark.parametrize("level", [1, 6])
def test_primed_chunks_decode_with_dictionary(gpu_ctx, oracle, level):
    # config 2 plan: deflate-raw, 64 KiB chunks, each primed with the preceding 32 KiB
    B = pkg("batch")
    data = make_text(600000, 31)
    r = B.deflate_batch(data, 65536, level, 0, B.MODE_INDEPENDENT, B.FLAG_PRIME)
    n_chunks = -(-len(data) // 65536)
    for i in range(n_chunks):
        lo = i * 65536
        chunk = data[lo: lo + 65536]
        dic = data[max(0, lo - 32768): lo] or None
        assert _decode_both(oracle, r.stream(i), 0, len(chunk), dic) == chunk, i


def test_ragged_chunks(gpu_ctx, oracle):
    B = pkg("batch")
    rng = np.random.default_rng(8)
    lens = np.concatenate([[0, 1, 2, 3, 258, 259, 70000, 0, 31, 32, 33], rng.integers(0, 5000, size=40)])
    off = np.zeros(lens.size + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    data = make_text(int(off[-1]), 9)
    r = B.deflate_batch(data, 0, 6, 2, B.MODE_INDEPENDENT, in_off=off)
    for i in range(lens.size):
        chunk = data[int(off[i]): int(off[i + 1])]
        assert _decode_both(oracle, r.stream(i), 2, len(chunk)) == chunk, (i, lens[i])


@pytest.mark.parametrize("level", [1, 6])
@pytest.mark.parametrize("prime", [False, True])
def test_many_tiny_ragged_chunks(gpu_ctx, oracle, level, prime):
    """Thousands of chunks of 0..300 bytes: one LZ77 segment holds > 100 of them, so chunk boundaries
    fall several to a 64-position parse hop (empty chunks, 1-byte chunks, boundaries at every offset)."""
    B = pkg("batch")
    rng = np.random.default_rng(12 + level)
    lens = rng.integers(0, 300, size=20000)
    lens[rng.integers(0, lens.size, 800)] = 0
    lens[rng.integers(0, lens.size, 50)] = rng.integers(300, 40000, 50)     # a few multi-block chunks in between
    off = np.zeros(lens.size + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    data = make_mixed(int(off[-1]), 10)
    r = B.deflate_batch(data, 0, level, 0, B.MODE_INDEPENDENT, flags=B.FLAG_PRIME if prime else 0, in_off=off)
    assert int(r.out_off[-1]) == len(r.data)
    for i in range(lens.size):
        lo, hi = int(off[i]), int(off[i + 1])
        d = zlib.decompressobj(-15, zdict=data[max(0, lo - 32768): lo]) if (prime and lo) else zlib.decompressobj(-15)
        assert d.decompress(r.stream(i)) + d.flush() == data[lo:hi] and d.eof, (i, lens[i])
    # the oracle on a sample (same verdicts as C zlib, slower)
    for i in range(0, lens.size, 997):
        lo, hi = int(off[i]), int(off[i + 1])
        ret, out, used, _ = oracle.inflate(r.stream(i), -15, hi - lo + 64, data[max(0, lo - 32768): lo] if (prime and lo) else None)
        assert ret == oracle.Z_STREAM_END and out == data[lo:hi]


def _kinds(n):
    rng = np.random.default_rng(5)
    pool = rng.integers(0, 256, (64, 3), dtype=np.uint8)
    sr = rng.integers(0, 256, n, dtype=np.uint8)
    for i in range(0, n - 9, 9):
        sr[i:i + 3] = pool[rng.integers(0, 64)]
    rows = b"".join(b'{"id":%d,"name":"user%d","score":%d,"tags":["a","b%d"]},' % (i, i * 7 % 1000, i * 13 % 97, i % 5) for i in range(n // 50))
    return {
        "text": make_text(n, 1), "mixed": make_mixed(n, 2),
        "dna": bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), n)),        # 4-letter alphabet: very long hash chains
        "short_repeats": sr.tobytes(),                                          # 3-byte matches only
        "counters": np.arange(n // 4, dtype=np.uint32).tobytes(),               # 3 of every 4 bytes repeat
        "floats": np.cumsum(rng.normal(0, 1, n // 8)).astype(np.float64).tobytes(),
        "json": rows[:n],
    }


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5, 6, 9])
def test_size_across_data_kinds(gpu_ctx, level):
    """Size against C zlib at the same level and with the same block plan (one stream, a block boundary
    after every 64 KiB chunk) on data the bench corpora do not cover.  Gate 3 % for every kind and level,
    no exceptions.  (Round 1 needed 5.5 % for JSON-like rows at levels 1-2: the engine inserts every position
    into the hash chains where deflate_fast skips the inside of long matches (deflate.ts:1310-1322), and so
    found many far 3-byte matches that cost more than literals; the greedy levels now apply deflate_slow's
    TOO_FAR rule, zs_lz77.cu search_position.)"""
    B = pkg("batch")
    n = 1 << 20
    for name, data in _kinds(n).items():
        r = B.deflate_batch(data, 65536, level, 1, B.MODE_STITCHED)
        assert zlib.decompress(r.data) == data
        co = zlib.compressobj(level)
        ref = 0
        for i in range(0, len(data), 65536):
            ref += len(co.compress(data[i:i + 65536])) + len(co.flush(zlib.Z_BLOCK))
        ref += len(co.flush())
        assert len(r.data) <= ref * RATIO_TOL + 64, (name, level, len(r.data), ref, len(r.data) / ref)


@pytest.mark.parametrize("level", [1, 5, 9])
def test_stitched_ragged_chunks_and_determinism(gpu_ctx, oracle, level):
    """One stream stitched from thousands of ragged chunks (with and without sync markers), twice: the
    bytes must not depend on how segments were scheduled."""
    B = pkg("batch")
    rng = np.random.default_rng(40 + level)
    lens = np.concatenate([rng.integers(0, 200, size=6000), rng.integers(900, 1100, size=300), [0, 0, 1, 959, 960, 961, 1919, 1920, 1921, 0]])
    rng.shuffle(lens)
    off = np.zeros(lens.size + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    data = make_mixed(int(off[-1]), 11)
    for flags in (0, B.FLAG_SYNC):
        r1 = B.deflate_batch(data, 0, level, 1, B.MODE_STITCHED, flags=flags, in_off=off)
        r2 = B.deflate_batch(data, 0, level, 1, B.MODE_STITCHED, flags=flags, in_off=off)
        assert r1.data == r2.data
        assert zlib.decompress(r1.data) == data
        ret, out, used, check = oracle.inflate(r1.data, 15, len(data) + 64)
        assert ret == oracle.Z_STREAM_END and out == data and used == len(r1.data) and check == zlib.adler32(data)


@pytest.mark.parametrize("level", [1, 6])
def test_compressed_size_within_tolerance(gpu_ctx, oracle, level):
    B = pkg("batch")
    report = {}
    for name, chunk in (("text", 65536), ("text", 262144), ("mixed", 262144)):
        data = make_text(2 << 20, 41) if name == "text" else make_mixed(3 << 20, 42)
        ref_stream = len(oracle.deflate(data, level, 1))                 # the reference, one continuous stream
        # the reference under the same chunk + 32 KiB dictionary plan (deflateSetDictionary + Z_SYNC_FLUSH)
        ref_plan = 0
        n_chunks = -(-len(data) // chunk)
        for i in range(n_chunks):
            lo = i * chunk
            last = i + 1 == n_chunks
            ref_plan += len(oracle.deflate(data[lo: lo + chunk], level, 0, data[max(0, lo - 32768): lo] or None,
                                           oracle.Z_FINISH if last else oracle.Z_SYNC_FLUSH))
        got = B.deflate_batch(data, chunk, level, 1, B.MODE_STITCHED).total_out_bytes
        report[(name, chunk)] = (got, ref_stream, ref_plan)
        assert got <= RATIO_TOL * ref_stream, (name, chunk, got, ref_stream)
        assert got <= RATIO_TOL * (ref_plan + 6), (name, chunk, got, ref_plan)
    print("sizes gpu / reference stream / reference chunk-plan:", report)


def test_device_api_large_roundtrip(gpu_ctx, oracle):
    # size-independent properties at a larger size: decode == input, checksum of checksums
    import torch
    B = pkg("batch")
    n = 48 << 20
    t = pkg("corpus").text_torch(n, torch.device("cuda:0"), seed=77)
    host = t.cpu().numpy().tobytes()
    res = B.deflate_batch_dev(t, 262144, 6, B.WRAP_ZLIB, B.MODE_STITCHED)
    rr = res.read_result()
    stream = res.out[: rr.total_out_bytes].cpu().numpy().tobytes()
    assert zlib.decompress(stream) == host
    assert rr.check == zlib.adler32(host)
    ratio = rr.total_out_bytes / n
    ref = len(zlib.compress(host[: 8 << 20], 6)) / (8 << 20)
    assert ratio <= RATIO_TOL * ref, (ratio, ref)
    # inflate the same stream on the GPU as one stream (single warp) -- bit exact
    d_in = res.out[: rr.total_out_bytes + 16]
    in_off = torch.tensor([0, rr.total_out_bytes], dtype=torch.int64, device="cuda")
    out_off = torch.tensor([0, n], dtype=torch.int64, device="cuda")
    inf = B.inflate_batch_dev(d_in, in_off, out_off, 15)
    torch.cuda.synchronize()
    assert int(inf.status[0]) == 1 and int(inf.out_len[0]) == n
    assert torch.equal(inf.out[:n], t)
    assert int(inf.checks[0]) & 0xFFFFFFFF == zlib.adler32(host)


def test_bit_concat(gpu_ctx, oracle):
    # K9: two parts of one stream produced separately, stitched at bit granularity
    import ctypes as C
    import torch
    B = pkg("batch")
    capi = pkg("capi")
    data = make_text(400000, 51)
    cut = 3 * 65536
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    a = B.deflate_batch_dev(t[:cut], 65536, 6, B.WRAP_RAW, B.MODE_STITCHED, B.FLAG_NOT_LAST)
    b = B.deflate_batch_dev(t[cut:], 65536, 6, B.WRAP_RAW, B.MODE_STITCHED, B.FLAG_NOT_FIRST, history=32768)
    ra, rb = a.read_result(), b.read_result()
    total_bits = ra.total_out_bits + rb.total_out_bits
    dst = torch.zeros((total_bits + 7) // 8 + 16, dtype=torch.uint8, device="cuda")
    ctx = B.default_context(0)
    lib = capi.load()
    ctx.check(lib.zs_bit_concat_dev(ctx.handle, C.c_void_p(dst.data_ptr()), 0, C.c_void_p(a.out.data_ptr()), ra.total_out_bits), "concat a")
    ctx.check(lib.zs_bit_concat_dev(ctx.handle, C.c_void_p(dst.data_ptr()), ra.total_out_bits, C.c_void_p(b.out.data_ptr()), rb.total_out_bits), "concat b")
    torch.cuda.synchronize()
    stream = dst[: (total_bits + 7) // 8].cpu().numpy().tobytes()
    assert _decode_both(oracle, stream, 0, len(data)) == data


def test_randomised_soak(gpu_ctx):
    """tools/soak.py for a few seconds: random data kind / size / chunking (down to 1-byte chunks) / level /
    strategy / wrapper / mode; every result decoded by C zlib, every call run twice (determinism).  This is
    the test that found the stale reads at the end of a range when segments are a few bytes long."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "soak.py"), "12", "7"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "soak ok" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
