# deflate throughput across chunk sizes / levels / modes (device-resident input)
import sys, os, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
dev = torch.device("cuda:0")
n = 512 << 20
ctx = B.default_context(0)
for name, gen in (("text", lambda: corpus.text_torch(n, dev, seed=5)), ("mixed", lambda: corpus.mixed_torch(n, dev))):
    t = gen()
    for chunk, lvl, wrap, mode, flags in ((4096, 6, B.WRAP_GZIP, B.MODE_INDEPENDENT, 0), (65536, 1, B.WRAP_RAW, B.MODE_INDEPENDENT, B.FLAG_PRIME),
                                          (65536, 6, B.WRAP_ZLIB, B.MODE_STITCHED, 0), (262144, 6, B.WRAP_ZLIB, B.MODE_STITCHED, 0),
                                          (262144, 9, B.WRAP_ZLIB, B.MODE_STITCHED, 0), (1 << 20, 6, B.WRAP_GZIP, B.MODE_INDEPENDENT, 0)):
        r = B.deflate_batch_dev(t, chunk, lvl, wrap, mode, flags, ctx=ctx)
        torch.cuda.synchronize()
        ctx.profile(True); ctx.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        B.deflate_batch_dev(t, chunk, lvl, wrap, mode, flags, ctx=ctx, reuse=r)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        prof = ctx.profile_read(); ctx.profile(False)
        rr = r.read_result()
        top = sorted(prof.items(), key=lambda kv: -kv[1][1])[:4]
        print(f"{name} chunk {chunk} L{lvl} wrap {wrap} mode {mode}: {n/ms/1e6:.2f} GB/s ratio {rr.total_out_bytes/n:.4f} blocks {rr.n_blocks}", {k: round(v[1], 2) for k, v in top})
    del t
