# host-buffer deflate (zs_deflate_batch / zs_deflate_part) from pinned memory: configs[1] and configs[2] shapes on one GPU
import sys, os, importlib, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
capi = importlib.import_module("zlib-streams-ts_b200.capi")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
lib = capi.load()
dev = torch.device("cuda:0")
ctx = B.default_context(0)
for name, gen, n, chunk, level, wrap, mode, flags in (
        ("configs[1] text L1 64K primed", lambda n: corpus.text_torch(n, dev, seed=0xC0FFEE), 1 << 30, 65536, 1, B.WRAP_RAW, B.MODE_INDEPENDENT, B.FLAG_PRIME),
        ("configs[2] mixed L6 256K stitched", lambda n: corpus.mixed_torch(n, dev, seed=0xB200), 1 << 30, 262144, 6, B.WRAP_ZLIB, B.MODE_STITCHED, 0)):
    data = gen(n)
    n_chunks = B.n_chunks_for(n, chunk)
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.copy_(data)
    cap = int(lib.zs_deflate_batch_bound(n, n_chunks, chunk, wrap, mode)) + 64
    h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    res = capi.DeflateResult()
    def step():
        rc = lib.zs_deflate_batch(ctx.handle, C.c_void_p(h_in.data_ptr()), n, None, n_chunks, chunk, level, wrap, mode, flags,
                                  C.c_void_p(h_out.data_ptr()), cap, None, None, None, C.byref(res))
        ctx.check(rc, "zs_deflate_batch")
    step(); step(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4): step()
    ms = 1e3 * (time.perf_counter() - t0) / 4
    r = B.deflate_batch_dev(data, chunk, level, wrap, mode, flags, ctx=ctx); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): B.deflate_batch_dev(data, chunk, level, wrap, mode, flags, ctx=ctx, reuse=r)
    e1.record(); torch.cuda.synchronize()
    dms = e0.elapsed_time(e1) / 3
    print(f"{name}: e2e {n/ms/1e6:.2f} GB/s ({ms:.1f} ms, {res.total_out_bytes} bytes out)  device {n/dms/1e6:.2f} GB/s ({dms:.1f} ms)  [ZS_SLICE_WAVES={os.environ.get('ZS_SLICE_WAVES')}]", flush=True)
    del data, h_in, h_out, r
