# inflate throughput of both kernels over (record size, record count)
import sys, os, importlib, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    B = importlib.import_module("zlib-streams-ts_b200.batch")
    corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
    dev = torch.device("cuda:0")
    for rec, nrec in ((4096, 16384), (4096, 32768), (4096, 65536), (4096, 131072), (16384, 16384), (16384, 32768), (65536, 4096), (65536, 16384)):
        n = rec * nrec
        t = corpus.text_torch(n, dev, seed=5)
        ioff = torch.arange(0, nrec + 1, dtype=torch.int64, device=dev) * rec
        r = B.deflate_batch_dev(t, rec, 6, B.WRAP_GZIP, B.MODE_INDEPENDENT, in_off=ioff, max_chunk=rec)
        torch.cuda.synchronize()
        inf = B.inflate_batch_dev(r.out, r.out_off, ioff, 31, out_capacity=n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            B.inflate_batch_dev(r.out, r.out_off, ioff, 31, reuse=inf)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        ok = bool((inf.status == 1).all().item()) and bool(torch.equal(inf.out[:n], t))
        print(f"  {nrec:7d} x {rec:6d}: {n/ms/1e6:6.2f} GB/s ok={ok}", flush=True)
        del t, r, inf
else:
    for name, var in (("warp per stream", "ZS_INFLATE_WARP"), ("thread per stream", "ZS_INFLATE_TPS")):
        print(name, flush=True)
        env = dict(os.environ); env[var] = "1"
        subprocess.run([sys.executable, __file__, "child"], env=env)
