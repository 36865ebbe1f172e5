# role cycles (ZS_LZ_PROF build) on the runs512 data at a given level
import sys, os, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
n = 64 << 20
rng = np.random.default_rng(1)
kind = sys.argv[1] if len(sys.argv) > 1 else "runs512"
if kind == "runs512":
    d = np.repeat(rng.integers(0, 256, size=n // 512, dtype=np.uint8), 512)
elif kind == "zeros":
    d = np.zeros(n, dtype=np.uint8)
t = torch.from_numpy(d).cuda()
for lvl in [int(x) for x in sys.argv[2:]] or [1, 6]:
    r = B.deflate_batch_dev(t, 262144, lvl, B.WRAP_ZLIB, B.MODE_STITCHED, 0)
    torch.cuda.synchronize()
    print(kind, "level", lvl, "ratio", r.read_result().total_out_bytes / n, flush=True)
