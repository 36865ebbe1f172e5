"""Policy exploration on the CPU: runs tools/lzmodel.c (a model of lz77_kernel's matching policy) over the
eight data kinds of tools/ratiocheck.py and prints, per kind and level, the modelled size against C zlib with
the same block plan and the modelled chain-walk work.  No GPU, no engine, no oracle.

usage: lzmodel.py [levels, e.g. 6 9] [-- key=value options of lzmodel ...]"""
import os
import subprocess
import sys
import tempfile
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from conftest import make_mixed, make_text  # noqa: E402

n = 4 << 20
rng = np.random.default_rng(5)


def dna():
    return bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), n))


def short_repeats():
    pool = rng.integers(0, 256, (64, 3), dtype=np.uint8)
    out = rng.integers(0, 256, n, dtype=np.uint8)
    for i in range(0, n - 9, 9):
        out[i:i + 3] = pool[rng.integers(0, 64)]
    return out.tobytes()


def json_like():
    rows = [b'{"id":%d,"name":"user%d","score":%d,"tags":["a","b%d"]},' % (i, i * 7 % 1000, i * 13 % 97, i % 5) for i in range(n // 50)]
    return b"".join(rows)[:n]


def source_code():
    parts = [("zlib-streams-ts_b200", "csrc", "zs_lz77.cu"), ("zlib-streams-ts_b200", "csrc", "zs_huff.cu"), ("SURVEY.md",), ("DESIGN.md",)]
    return b"".join(open(os.path.join(ROOT, *p), "rb").read() for p in parts)


KINDS = {"text": lambda: make_text(n, 1), "mixed": lambda: make_mixed(n, 2), "dna": dna, "short_repeats": short_repeats,
         "counters": lambda: np.arange(n // 4, dtype=np.uint32).tobytes(),
         "floats": lambda: np.cumsum(rng.normal(0, 1, n // 8)).astype(np.float64).tobytes(), "json": json_like, "source": source_code}


def main():
    args = sys.argv[1:]
    opts = []
    if "--" in args:
        opts = args[args.index("--") + 1:]
        args = args[: args.index("--")]
    levels = [int(x) for x in args] or [1, 6, 9]
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "lzmodel")
        subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tools", "lzmodel.c")], check=True)
        worst = 0.0
        for name, gen in KINDS.items():
            data = gen()
            path = os.path.join(tmp, name + ".bin")
            open(path, "wb").write(data)
            row = []
            for lvl in levels:
                co = zlib.compressobj(lvl, 8, -15)
                ref = 0
                for i in range(0, len(data), 65536):
                    ref += len(co.compress(data[i:i + 65536])) + len(co.flush(zlib.Z_BLOCK))
                ref += len(co.flush())
                out = subprocess.run([exe, path, str(lvl), "65536", "16", *opts], capture_output=True, text=True, check=True).stdout.split("\n")
                size = int(out[0].split()[3])
                heaviest = float(out[3].split()[-1])
                walk = float(out[1].split("(")[2].split()[0])
                row.append(f"L{lvl} {size / ref:6.4f} walk {walk:6.2f} step {heaviest:7.2f}")
                worst = max(worst, size / ref)
            print(f"{name:14s} {len(data):8d} B  model/zlib, longest walk per batch, heaviest batch per step: " + " | ".join(row), flush=True)
        print("worst", round(worst, 4))


if __name__ == "__main__":
    main()
