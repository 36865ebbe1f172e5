# level-6 / level-9 deflate throughput per kind of data (the five ingredients of the mixed corpus)
import sys, os, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
dev = torch.device("cuda:0")
n = 128 << 20
rng = np.random.default_rng(1)
def holes():
    blk = rng.integers(0, 256, size=4096, dtype=np.uint8)
    rep = np.tile(blk, n // 4096 + 1)[:n].copy()
    h = rng.integers(0, n, size=n // 2048)
    rep[h] = rng.integers(0, 256, size=h.size, dtype=np.uint8)
    return rep
kinds = {
    "text": lambda: corpus.text_numpy(n, 5),
    "ramp251": lambda: (np.arange(n, dtype=np.int64) % 251).astype(np.uint8),
    "random": lambda: rng.integers(0, 256, size=n, dtype=np.uint8),
    "zeros": lambda: np.zeros(n, dtype=np.uint8),
    "runs512": lambda: np.repeat(rng.integers(0, 256, size=n // 512, dtype=np.uint8), 512),
    "repeat4k+holes": holes,
}
ctx = B.default_context(0)
levels = [int(x) for x in sys.argv[1:]] or [1, 6, 9]
for name, gen in kinds.items():
    t = torch.from_numpy(gen()).to(dev)
    for lvl in levels:
        r = B.deflate_batch_dev(t, 262144, lvl, B.WRAP_ZLIB, B.MODE_STITCHED, 0, ctx=ctx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        B.deflate_batch_dev(t, 262144, lvl, B.WRAP_ZLIB, B.MODE_STITCHED, 0, ctx=ctx, reuse=r)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{name:16s} L{lvl}: {n/ms/1e6:7.2f} GB/s  ratio {r.read_result().total_out_bytes/n:.4f}", flush=True)
    del t
