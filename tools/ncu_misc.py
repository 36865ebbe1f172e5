# small driver for ncu captures of the inflate and checksum kernels
import sys, os, importlib, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
dev = torch.device("cuda:0")
n = 64 << 20
t = corpus.text_torch(n, dev, seed=5)
# checksum of 64 MiB in 64 KiB segments
off = torch.arange(0, n + 1, 65536, dtype=torch.int64, device=dev)
B.checksum_batch_dev(t, off, 1); B.checksum_batch_dev(t, off, 0)
# inflate: 16384 gzip records of 4 KiB (config 4 shape), compressed on the GPU with level 6
rec = 4096
nrec = n // rec
ioff = torch.arange(0, nrec + 1, dtype=torch.int64, device=dev) * rec
r = B.deflate_batch_dev(t, rec, 6, B.WRAP_GZIP, B.MODE_INDEPENDENT, in_off=ioff, max_chunk=rec)
torch.cuda.synchronize()
inf = B.inflate_batch_dev(r.out, r.out_off, ioff, 31, out_capacity=n)
torch.cuda.synchronize()
ok = bool((inf.status == 1).all().item()) and bool(torch.equal(inf.out[:n], t))
print("ok", ok, int(r.read_result().total_out_bytes))
