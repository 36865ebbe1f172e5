import sys, os, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
n = 128 << 20
t = corpus.text_torch(n, torch.device("cuda:0"), seed=5)
for lvl in (1, 6):
    r = B.deflate_batch_dev(t, 65536, lvl, B.WRAP_RAW, B.MODE_INDEPENDENT, B.FLAG_PRIME)
    torch.cuda.synchronize()
    print('level', lvl, 'ratio', r.read_result().total_out_bytes / n)
