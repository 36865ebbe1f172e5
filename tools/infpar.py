"""Throughput of the segment-parallel decoder of one stream (zs_inflate_stream_dev): the engine's own STITCHED +
SYNC deflate of a 1 GiB corpus, and a C zlib stream with flush points.  usage: infpar.py [MiB] [chunk KiB]"""
import ctypes as C, importlib, os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
capi = importlib.import_module("zlib-streams-ts_b200.capi")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
lib = capi.load()
dev = torch.device("cuda:0")
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
chunk = (int(sys.argv[2]) if len(sys.argv) > 2 else 64) << 10
n = mib << 20
ctx = B.default_context(0)
res = torch.zeros(4, dtype=torch.int64, device=dev)


def run(d_in, in_len, wbits, d_out, label, reps=3):
    def once():
        rc = lib.zs_inflate_stream_dev(ctx.handle, d_in.data_ptr(), in_len, wbits, d_out.data_ptr(), d_out.numel(), res.data_ptr(),
                                       res.data_ptr() + 8, res.data_ptr() + 16, res.data_ptr() + 24, None, 0)
        ctx.check(rc, "zs_inflate_stream_dev")
    once(); torch.cuda.synchronize()
    ctx.profile(True); ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        once()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    prof = ctx.profile_read(); ctx.profile(False)
    out_len, in_used, check, status = (int(x) for x in res.cpu())
    print(f"{label}: status {status & 0xffffffff} out {out_len} in_used {in_used}/{in_len}  {out_len / ms / 1e6:.2f} GB/s of output ({ms:.1f} ms)",
          {k: round(v[1] / reps, 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}, flush=True)
    return out_len, status & 0xffffffff


for name, gen in (("text", lambda: corpus.text_torch(n, dev, seed=5)), ("mixed", lambda: corpus.mixed_torch(n, dev))):
    t = gen()
    for level in (1, 6):
        r = B.deflate_batch_dev(t, chunk, level, B.WRAP_ZLIB, B.MODE_STITCHED, B.FLAG_SYNC, ctx=ctx)
        torch.cuda.synchronize()
        comp = int(r.read_result().total_out_bytes)
        d_out = torch.empty(n + 64, dtype=torch.uint8, device=dev)
        out_len, status = run(r.out, comp, 15, d_out, f"{name} {mib} MiB own stream L{level} {chunk >> 10} KiB chunks + sync (ratio {comp / n:.3f})")
        print("   bit exact:", status == 1 and out_len == n and bool(torch.equal(d_out[:n], t)), flush=True)
        del r, d_out
    del t
# a C zlib stream with a sync flush every 128 KiB (what pigz writes), 256 MiB
m = min(n, 256 << 20)
host = corpus.text_numpy(m, 3).tobytes()
co = zlib.compressobj(6)
parts = []
t0 = time.time()
for i in range(0, m, 131072):
    parts.append(co.compress(host[i:i + 131072])); parts.append(co.flush(zlib.Z_SYNC_FLUSH))
parts.append(co.flush())
z = b"".join(parts)
print(f"C zlib deflate of {m >> 20} MiB with sync flushes: {time.time() - t0:.1f} s", flush=True)
t0 = time.time(); zlib.decompress(z); print(f"C zlib inflate, one core: {m / (time.time() - t0) / 1e9:.3f} GB/s", flush=True)
d_in = torch.frombuffer(bytearray(z + bytes(8)), dtype=torch.uint8).to(dev)
d_out = torch.empty(m + 64, dtype=torch.uint8, device=dev)
out_len, status = run(d_in, len(z), 15, d_out, f"C zlib stream, sync flush every 128 KiB, {m >> 20} MiB")
print("   bit exact:", status == 1 and bytes(d_out[:m].cpu().numpy()) == host, flush=True)
# the same data written in one go (no flush points): cut at speculatively located dynamic block headers
t0 = time.time(); z2 = zlib.compress(host, 6); print(f"C zlib deflate of {m >> 20} MiB in one go: {time.time() - t0:.1f} s", flush=True)
d_in2 = torch.frombuffer(bytearray(z2 + bytes(8)), dtype=torch.uint8).to(dev)
out_len, status = run(d_in2, len(z2), 15, d_out, f"C zlib stream without flush points, {m >> 20} MiB")
print("   bit exact:", status == 1 and bytes(d_out[:m].cpu().numpy()) == host, flush=True)
