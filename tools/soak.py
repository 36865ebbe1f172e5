"""Randomised round-trip soak of the deflate engine: random data kind / size / chunking / level / strategy /
wrapper / mode, every result decoded (C zlib on the host) and compared, every call run twice (determinism).
usage: soak.py [seconds] [seed]"""
import sys, os, importlib, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import make_mixed, make_text
B = importlib.import_module("zlib-streams-ts_b200.batch")
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
WB = {0: -15, 1: 15, 2: 31}

def make(kind, n):
    if kind == 0: return make_text(n, int(rng.integers(1 << 30)))
    if kind == 1: return make_mixed(n, int(rng.integers(1 << 30)))
    if kind == 2: return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if kind == 3: return bytes(n)
    if kind == 4: return np.repeat(rng.integers(0, 256, n // 37 + 1, dtype=np.uint8), 37)[:n].tobytes()
    if kind == 5: return b"".join(b'{"id":%d,"v":"%d"},' % (i, i * 7919 % 1000) for i in range(n // 18 + 1))[:n]
    return (np.arange(n) % 251).astype(np.uint8).tobytes()

t0 = time.time(); it = 0; total = 0
while time.time() - t0 < budget:
    it += 1
    n = int(rng.choice([0, 1, 2, 3, 100, 5000, 70000, 300000, 1 << 20, 3 << 20, 8 << 20], p=[.02, .02, .02, .02, .1, .15, .2, .2, .15, .08, .04]))
    data = make(int(rng.integers(0, 7)), n)
    level = int(rng.integers(0, 10)); strategy = int(rng.choice([0, 0, 0, 1, 2, 3, 4])); wrap = int(rng.integers(0, 3))
    mode = int(rng.integers(0, 2))
    flags = B.flag_strategy(strategy)
    prime = False
    if mode == B.MODE_INDEPENDENT and wrap == 0 and rng.random() < 0.5: flags |= B.FLAG_PRIME; prime = True
    if mode == B.MODE_STITCHED and rng.random() < 0.3: flags |= B.FLAG_SYNC
    if rng.random() < 0.35 and n > 0:
        k = int(rng.integers(1, 400))
        cuts = np.sort(rng.integers(0, n + 1, k))
        off = np.concatenate([[0], cuts, [n]]).astype(np.uint64)
        chunk = 0
    else:
        off = None
        chunk = int(rng.choice([1, 7, 100, 1000, 4096, 65536, 262144, 1 << 20]))
        if n // max(chunk, 1) > 200000: chunk = 4096
    desc = (it, n, level, strategy, wrap, mode, flags, chunk, None if off is None else off.size - 1)
    r1 = B.deflate_batch(data, chunk, level, wrap, mode, flags=flags, in_off=off)
    r2 = B.deflate_batch(data, chunk, level, wrap, mode, flags=flags, in_off=off)
    assert r1.data == r2.data, ("nondeterministic", desc)
    if mode == B.MODE_STITCHED:
        d = zlib.decompressobj(WB[wrap])
        out = d.decompress(r1.data) + d.flush()
        assert out == data and d.eof, ("stitched decode", desc)
    else:
        offs = off if off is not None else np.array(list(range(0, n, chunk)) + [n] if n else [0, 0], dtype=np.uint64)
        nch = offs.size - 1
        idx = range(nch) if nch <= 3000 else rng.choice(nch, 3000, replace=False)
        for i in idx:
            lo, hi = int(offs[i]), int(offs[i + 1])
            zd = data[max(0, lo - 32768): lo] if (prime and lo) else None
            d = zlib.decompressobj(WB[wrap], zdict=zd) if zd else zlib.decompressobj(WB[wrap])
            out = d.decompress(r1.stream(int(i))) + d.flush()
            assert out == data[lo:hi] and d.eof, ("independent decode", desc, int(i))
    total += n
print(f"soak ok: {it} calls, {total / 1e6:.1f} MB, seed {seed}, {time.time() - t0:.0f} s")
