import sys, os, importlib, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import make_mixed
B = importlib.import_module("zlib-streams-ts_b200.batch")
level = int(sys.argv[1]) if len(sys.argv) > 1 else 1
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(12 + level)
lens = rng.integers(0, 300, size=20000) * scale
lens[rng.integers(0, lens.size, 800)] = 0
lens[rng.integers(0, lens.size, 50)] = rng.integers(300, 40000, 50)
off = np.zeros(lens.size + 1, dtype=np.uint64)
np.cumsum(lens, out=off[1:])
data = make_mixed(int(off[-1]), 10)
r = B.deflate_batch(data, 0, level, 0, B.MODE_INDEPENDENT, flags=B.FLAG_PRIME, in_off=off)
bad = []
for i in range(lens.size):
    lo, hi = int(off[i]), int(off[i + 1])
    d = zlib.decompressobj(-15, zdict=data[max(0, lo - 32768): lo]) if lo else zlib.decompressobj(-15)
    try:
        out = d.decompress(r.stream(i)) + d.flush()
    except Exception as e:
        out = repr(e).encode()
    if out != data[lo:hi]:
        bad.append(i)
print("level", level, "scale", scale, "grid", os.environ.get("ZS_LZ_GRID"), "bad chunks:", len(bad), bad[:12], "idx in segment:", sorted(set(b % 16 for b in bad)))
sys.exit(0)
for i in bad[:6]:
    for j in range(i - 2, i + 3):
        lo, hi = int(off[j]), int(off[j + 1])
        d = zlib.decompressobj(-15, zdict=data[max(0, lo - 32768): lo]) if lo else zlib.decompressobj(-15)
        try:
            out = d.decompress(r.stream(j)) + d.flush()
        except Exception as e:
            out = repr(e).encode()
        if j == i: print("   extra bytes", out[hi - lo:][:24], "next data", data[hi:hi + 24], "n-rel-to-64:", (hi - int(off[(j // 16) * 16])) % 64)
        print(j, "len", lens[j], "off", lo, "(seg", j // 16, "idx", j % 16, ") range-rel", lo - int(off[(j // 16) * 16]), "got", len(out), "want", hi - lo, "tail ok" if out == data[lo:hi] else ("got[:8]=%r want[:8]=%r" % (out[:8], data[lo:hi][:8])))
