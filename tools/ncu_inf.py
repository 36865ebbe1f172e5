# driver for ncu captures of the warp-per-stream inflate kernel: N MiB of text in 64 KiB raw streams
import sys, os, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
dev = torch.device("cuda:0")
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 256) << 20
rec = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
t = corpus.text_torch(n, dev, seed=5)
nrec = n // rec
ioff = torch.arange(0, nrec + 1, dtype=torch.int64, device=dev) * rec
r = B.deflate_batch_dev(t, rec, 6, B.WRAP_RAW, B.MODE_INDEPENDENT, in_off=ioff, max_chunk=rec)
torch.cuda.synchronize()
inf = B.inflate_batch_dev(r.out, r.out_off, ioff, -15, out_capacity=n)
torch.cuda.synchronize()
print("ok", bool((inf.status == 1).all().item()) and bool(torch.equal(inf.out[:n], t)))
