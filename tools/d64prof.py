# deflate64 batch inflate throughput (configs[4]): the reference's fixtures replicated to ~1 GiB of output
import sys, os, json, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
meta = json.load(open(os.path.join(ROOT, "tests", "golden", "deflate64_fixtures.json")))
blob = open(os.path.join(ROOT, "tests", "golden", "deflate64_fixtures.bin"), "rb").read()
fx = meta["fixtures"] if "fixtures" in meta else meta
dev = torch.device("cuda:0")
for name, reps in (("100k_lines.deflate64", 512), ("payload_64k.deflate64", 16384), ("zeros_100k.deflate64", 8192)):
    f = next(x for x in fx if x["name"] == name)
    z = blob[f["offset"]: f["offset"] + f["length"]]
    zpad = z + bytes((-len(z)) % 8)                      # keep every stream 8-byte aligned
    din = torch.from_numpy(np.frombuffer(zpad * reps, dtype=np.uint8).copy()).to(dev)
    ioff = torch.arange(0, reps + 1, dtype=torch.int64, device=dev) * len(zpad)
    cap = (f["out_len"] + 15) & ~15
    ooff = torch.arange(0, reps + 1, dtype=torch.int64, device=dev) * cap
    inf = B.inflate_batch_dev(din, ioff, ooff, -16, out_capacity=cap * reps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        B.inflate_batch_dev(din, ioff, ooff, -16, reuse=inf)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ok = bool((inf.status == 1).all().item()) and bool((inf.out_len == f["out_len"]).all().item())
    total = f["out_len"] * reps
    print(f"{name:26s} x {reps:6d}: ok={ok}  {total / ms / 1e6:7.2f} GB/s of output ({ms:.2f} ms, {total / 2**30:.2f} GiB)", flush=True)
