"""Randomised soak of inflate (batch, both kernels) and of the z_stream API against the oracle / C zlib.
usage: soak_inflate.py [seconds] [seed]"""
import sys, os, importlib, time, zlib, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

def make(rng, kind, n):
    from conftest import make_mixed, make_text
    if kind == 0: return make_text(n, int(rng.integers(1 << 30)))
    if kind == 1: return make_mixed(n, int(rng.integers(1 << 30)))
    if kind == 2: return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if kind == 3: return bytes(n)
    return np.repeat(rng.integers(0, 256, n // 37 + 1, dtype=np.uint8), 37)[:n].tobytes()

def batch_child(budget, seed):
    from oracle import oracle as O
    B = importlib.import_module("zlib-streams-ts_b200.batch")
    rng = np.random.default_rng(seed)
    t0 = time.time(); it = 0; ns = 0
    while time.time() - t0 < budget:
        it += 1
        wb = int(rng.choice([15, 31, -15, 47, -16]))
        count = int(rng.choice([1, 3, 40, 200]))
        streams, caps, raws = [], [], []
        for _ in range(count):
            n = int(rng.choice([0, 1, 50, 3000, 70000, 400000], p=[.05, .05, .2, .35, .25, .1]))
            data = make(rng, int(rng.integers(0, 5)), n)
            if wb == -16:
                z = O.deflate64_encode(data, int(rng.choice([258, 65538])))
            else:
                zwb = {15: 15, 31: 31, -15: -15, 47: int(rng.choice([15, 31]))}[wb]
                co = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, zwb, 8, int(rng.choice([0, 0, 1, 2, 3, 4])))
                z = co.compress(data) + co.flush()
            r = rng.random()
            if r < 0.15 and len(z) > 2:
                z = z[: int(rng.integers(1, len(z)))]
            elif r < 0.3 and len(z) > 2:
                b = bytearray(z); b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8)); z = bytes(b)
            elif r < 0.4:
                z = z + b"tail"
            cap = n + 16 if rng.random() < 0.85 else max(0, n - int(rng.integers(0, n + 1)))
            streams.append(z); caps.append(cap); raws.append(data)
        res = B.inflate_batch(streams, wb, caps)
        for i, z in enumerate(streams):
            ret, out, used, check = O.inflate(z, wb, caps[i])
            assert int(res.status[i]) == ret, ("status", seed, it, i, wb, int(res.status[i]), ret, res.message(i))
            assert res.output(i) == out, ("output", seed, it, i, wb, len(res.output(i)), len(out))
            if ret == O.Z_STREAM_END:
                assert int(res.in_used[i]) == used, ("in_used", seed, it, i, wb)
        ns += count
    print(f"batch soak ok ({os.environ.get('ZS_INFLATE_TPS') and 'thread' or 'warp'} kernel): {it} calls, {ns} streams, seed {seed}")

def api_child(budget, seed):
    Z = importlib.import_module("zlib-streams-ts_b200.zlib_api")
    rng = np.random.default_rng(seed)
    t0 = time.time(); it = 0
    while time.time() - t0 < budget:
        it += 1
        n = int(rng.choice([0, 1, 100, 5000, 70000, 600000], p=[.05, .05, .2, .3, .25, .15]))
        data = make(rng, int(rng.integers(0, 5)), n)
        wbits = int(rng.choice([15, 31, -15]))
        level = int(rng.integers(0, 10)); strategy = int(rng.choice([0, 0, 0, 1, 2, 3, 4]))
        s = Z.createDeflateStream()
        assert Z.deflateInit2_(s, level, 8, wbits, 8, strategy) == Z.Z_OK
        out = bytearray(); pos = 0
        while pos < n:
            piece = data[pos: pos + int(rng.choice([1, 13, 1000, 32768, 200000]))]
            flush = int(rng.choice([Z.Z_NO_FLUSH] * 6 + [Z.Z_SYNC_FLUSH, Z.Z_FULL_FLUSH, Z.Z_PARTIAL_FLUSH, Z.Z_BLOCK]))
            s.next_in, s.next_in_index, s.avail_in = piece, 0, len(piece)
            while True:
                cap = int(rng.choice([1, 64, 4096, 65536]))
                buf = bytearray(cap)
                s.next_out, s.next_out_index, s.avail_out = buf, 0, cap
                r = Z.deflate(s, flush)
                out += buf[: s.next_out_index]
                assert r in (Z.Z_OK, Z.Z_BUF_ERROR), (seed, it, r)
                if s.avail_in == 0 and s.avail_out != 0:
                    break
            pos += len(piece)
        while True:
            cap = int(rng.choice([1, 64, 4096, 65536]))
            buf = bytearray(cap)
            s.next_in, s.next_in_index, s.avail_in = b"", 0, 0
            s.next_out, s.next_out_index, s.avail_out = buf, 0, cap
            r = Z.deflate(s, Z.Z_FINISH)
            out += buf[: s.next_out_index]
            if r == Z.Z_STREAM_END:
                break
            assert r == Z.Z_OK, (seed, it, r)
        assert Z.deflateEnd(s) == Z.Z_OK
        d = zlib.decompressobj(wbits)
        assert d.decompress(bytes(out)) + d.flush() == data and d.eof, ("deflate stream", seed, it, n, wbits, level, strategy)
        # and back through inflate(), from C zlib's stream of the same data, in random pieces
        co = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, wbits)
        z = co.compress(data) + co.flush() + b"XYZ"
        t = Z.createInflateStream()
        assert Z.inflateInit2_(t, wbits) == Z.Z_OK
        got = bytearray(); pos = 0; r = Z.Z_OK
        while r != Z.Z_STREAM_END:
            piece = z[pos: pos + int(rng.choice([1, 7, 500, 32768, 300000]))]
            t.next_in, t.next_in_index, t.avail_in = piece, 0, len(piece)
            cap = int(rng.choice([1, 100, 65536, 1 << 20]))
            buf = bytearray(cap)
            t.next_out, t.next_out_index, t.avail_out = buf, 0, cap
            r = Z.inflate(t, int(rng.choice([Z.Z_NO_FLUSH, Z.Z_NO_FLUSH, Z.Z_SYNC_FLUSH])))
            assert r in (Z.Z_OK, Z.Z_STREAM_END, Z.Z_BUF_ERROR), (seed, it, r, t.msg)
            got += buf[: t.next_out_index]
            pos += len(piece) - t.avail_in
            assert pos <= len(z)
        # total_in is always exact; the bytes behind the end come back through avail_in when the stream is shorter than
        # the shim's every-call threshold (64 KiB of input) or the caller flushes -- a longer stream is decoded in
        # batches and may find its end one call after the input arrived ("Streaming inflate")
        assert bytes(got) == data and t.total_in == len(z) - 3, ("inflate stream", seed, it, n, wbits, t.total_in, len(z))
        assert pos == len(z) - 3 or (len(z) >= (64 << 10) and pos <= len(z)), ("inflate stream hand-back", seed, it, n, wbits, pos, len(z))
        assert Z.inflateEnd(t) == Z.Z_OK
    print(f"api soak ok: {it} streams each way, seed {seed}")

if __name__ == "__main__":
    if len(sys.argv) > 3:
        (batch_child if sys.argv[3] == "batch" else api_child)(float(sys.argv[1]), int(sys.argv[2]))
        sys.exit(0)
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rc = 0
    for mode, env in (("batch", {"ZS_INFLATE_WARP": "1"}), ("batch", {"ZS_INFLATE_TPS": "1"}), ("api", {})):
        e = dict(os.environ); e.update(env)
        rc |= subprocess.run([sys.executable, __file__, str(budget / 3), str(seed), mode], env=e).returncode
    sys.exit(rc)
