import sys, os, importlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
dev = torch.device("cuda:0")
n = 512 << 20
t = corpus.text_torch(n, dev, seed=5)
ctx = B.default_context(0)
for rec, wrap, wb in ((4096, B.WRAP_GZIP, 31), (65536, B.WRAP_RAW, -15), (1 << 20, B.WRAP_ZLIB, 15)):
    nrec = n // rec
    ioff = torch.arange(0, nrec + 1, dtype=torch.int64, device=dev) * rec
    r = B.deflate_batch_dev(t, rec, 6, wrap, B.MODE_INDEPENDENT, in_off=ioff, max_chunk=rec)
    torch.cuda.synchronize()
    comp = int(r.read_result().total_out_bytes)
    inf = B.inflate_batch_dev(r.out, r.out_off, ioff, wb, out_capacity=n)
    torch.cuda.synchronize()
    ctx.profile(True); ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        B.inflate_batch_dev(r.out, r.out_off, ioff, wb, reuse=inf)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    prof = ctx.profile_read(); ctx.profile(False)
    ok = bool((inf.status == 1).all().item()) and bool(torch.equal(inf.out[:n], t))
    print(f"records {nrec} x {rec}: ok={ok} comp={comp/n:.3f} inflate {n/ms/1e6:.2f} GB/s out ({ms:.2f} ms)", {k: round(v[1]/3, 3) for k, v in prof.items()})
