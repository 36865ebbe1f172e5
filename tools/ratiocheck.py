# compressed size against C zlib at the same level over several kinds of data (stitched 64 KiB chunks, zlib wrapper)
import sys, os, importlib, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import make_mixed, make_text
B = importlib.import_module("zlib-streams-ts_b200.batch")
rng = np.random.default_rng(5)
n = 4 << 20
def dna(): return bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), n))
def short_repeats():
    # random bytes where every third triple repeats one of 64 recent triples: many 3-byte matches, few longer ones
    pool = rng.integers(0, 256, (64, 3), dtype=np.uint8)
    out = rng.integers(0, 256, n, dtype=np.uint8)
    for i in range(0, n - 9, 9):
        out[i:i + 3] = pool[rng.integers(0, 64)]
    return out.tobytes()
def counters(): return np.arange(n // 4, dtype=np.uint32).tobytes()
def floats(): return np.cumsum(rng.normal(0, 1, n // 8)).astype(np.float64).tobytes()
def json_like():
    rows = [b'{"id":%d,"name":"user%d","score":%d,"tags":["a","b%d"]},' % (i, i * 7 % 1000, i * 13 % 97, i % 5) for i in range(n // 50)]
    return b"".join(rows)[:n]
def source_code():
    txt = open(os.path.join(ROOT, "zlib-streams-ts_b200", "csrc", "zs_lz77.cu"), "rb").read()
    return (txt * (n // len(txt) + 1))[:n] if False else (txt + open(os.path.join(ROOT, "zlib-streams-ts_b200", "csrc", "zs_huff.cu"), "rb").read() + open(os.path.join(ROOT, "SURVEY.md"), "rb").read() + open(os.path.join(ROOT, "DESIGN.md"), "rb").read())
kinds = {"text": lambda: make_text(n, 1), "mixed": lambda: make_mixed(n, 2), "dna": dna, "short_repeats": short_repeats, "counters": counters,
         "floats": floats, "json": json_like, "source": source_code}
levels = [int(x) for x in sys.argv[1:]] or [1, 3, 6, 9]
worst = 0
for name, gen in kinds.items():
    data = gen()
    row = []
    for lvl in levels:
        r = B.deflate_batch(data, 65536, lvl, 1, B.MODE_STITCHED)
        assert zlib.decompress(r.data) == data
        # the reference with the same plan: one stream, a block boundary (Z_BLOCK) after every 64 KiB chunk
        co = zlib.compressobj(lvl)
        ref = 0
        for i in range(0, len(data), 65536):
            ref += len(co.compress(data[i:i + 65536])) + len(co.flush(zlib.Z_BLOCK))
        ref += len(co.flush())
        one = len(zlib.compress(data, lvl))
        row.append(f"L{lvl} {len(r.data) / ref:6.4f} ({len(r.data) / one:6.4f})")
        worst = max(worst, len(r.data) / ref)
    print(f"{name:14s} {len(data):8d} B  gpu/zlib same plan (gpu/zlib one shot): " + "  ".join(row), flush=True)
print("worst", round(worst, 4))
