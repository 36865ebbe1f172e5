# driver for ncu captures of the kernels around lz77: checksums on 1 GiB (64 KiB segments), the Huffman stage
# kernels of a 256 MiB level-1 deflate, and the thread-per-stream inflate on 131072 gzip records of 4 KiB
import sys, os, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
dev = torch.device("cuda:0")
n = 1 << 30
t = corpus.text_torch(n, dev, seed=5)
off = torch.arange(0, n + 1, 65536, dtype=torch.int64, device=dev)
B.checksum_batch_dev(t, off, 1); B.checksum_batch_dev(t, off, 0)
torch.cuda.synchronize()
m = 256 << 20
r1 = B.deflate_batch_dev(t[:m], 65536, 1, B.WRAP_RAW, B.MODE_INDEPENDENT, B.FLAG_PRIME)
torch.cuda.synchronize()
# inflate: gzip records of 4 KiB (configs[3] shape), compressed on the GPU with level 6
rec, m2 = 4096, 512 << 20
nrec = m2 // rec
ioff = torch.arange(0, nrec + 1, dtype=torch.int64, device=dev) * rec
r = B.deflate_batch_dev(t[:m2], rec, 6, B.WRAP_GZIP, B.MODE_INDEPENDENT, in_off=ioff, max_chunk=rec)
torch.cuda.synchronize()
inf = B.inflate_batch_dev(r.out, r.out_off, ioff, 31, out_capacity=m2)
torch.cuda.synchronize()
ok = bool((inf.status == 1).all().item()) and bool(torch.equal(inf.out[:m2], t[:m2]))
print("ok", ok, int(r.read_result().total_out_bytes))
