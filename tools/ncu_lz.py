# small driver for ncu captures: one deflate call per level on a 32 MiB text corpus
import sys, os, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 32 << 20
lvl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t = corpus.text_torch(n, torch.device("cuda:0"), seed=5)
r = B.deflate_batch_dev(t, 65536, lvl, B.WRAP_RAW, B.MODE_INDEPENDENT, B.FLAG_PRIME)
torch.cuda.synchronize()
print("ok", r.read_result().total_out_bytes)
