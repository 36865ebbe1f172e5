# one-off size check beyond 2^32 bytes: 5 GiB in one call (stitched zlib level 6 with 256 KiB chunks = the
# shape of configs[2], and raw level 1 with primed 64 KiB chunks), decoded again on the device and compared
import sys, os, importlib, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
dev = torch.device("cuda:0")
n = 5 << 30
base = corpus.mixed_torch(1 << 30, dev)
data = base.repeat(5)[:n].contiguous()
del base
ctx = B.default_context(0)
# 1) independent primed chunks -> inflate batch with dictionaries
chunk = 65536
nch = n // chunk
r = B.deflate_batch_dev(data, chunk, 1, B.WRAP_RAW, B.MODE_INDEPENDENT, B.FLAG_PRIME, ctx=ctx)
torch.cuda.synchronize()
rr = r.read_result()
print("L1 primed: out bytes", rr.total_out_bytes, "ratio", rr.total_out_bytes / n, "blocks", rr.n_blocks, flush=True)
off = torch.arange(0, nch + 1, dtype=torch.int64, device=dev) * chunk
starts = off[:-1]
rng = torch.stack([torch.clamp(starts - 32768, min=0), starts], 1).reshape(-1).contiguous()
inf = B.inflate_batch_dev(r.out, r.out_off, off, -15, d_dict=data, dict_rng=rng, out_capacity=n, ctx=ctx)
torch.cuda.synchronize()
print("  inflate ok:", bool((inf.status == 1).all().item()), "equal:", bool(torch.equal(inf.out[:n], data)), flush=True)
del r, inf
torch.cuda.empty_cache()
# 2) one stitched zlib stream of 5 GiB: check adler32 and decode the head and the tail region on the host
r = B.deflate_batch_dev(data, 262144, 6, B.WRAP_ZLIB, B.MODE_STITCHED, 0, ctx=ctx)
torch.cuda.synchronize()
rr = r.read_result()
print("L6 stitched: out bytes", rr.total_out_bytes, "ratio", rr.total_out_bytes / n, flush=True)
stream = r.out[: rr.total_out_bytes].cpu().numpy().tobytes()
d = zlib.decompressobj()
ok = True
pos = 0
step = 64 << 20
adler = 1
for i in range(0, len(stream), step):
    out = d.decompress(stream[i:i + step])
    ref = data[pos: pos + len(out)].cpu().numpy().tobytes()
    ok &= out == ref
    pos += len(out)
out = d.flush(); ok &= out == data[pos: pos + len(out)].cpu().numpy().tobytes(); pos += len(out)
print("  zlib decode ok:", ok and d.eof and pos == n, "check equal:", rr.check == int.from_bytes(stream[-4:], "big"), flush=True)
