# repro driver for lz77 variants: primed independent chunks / stitched, several sizes, level from argv
import sys, os, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sizes = [int(x) for x in sys.argv[2:]] or [4, 32, 128]
dev = torch.device("cuda:0")
for mib in sizes:
    t = corpus.text_torch(mib << 20, dev, seed=5)
    for name, mode, flags, wrap in (("stitched", B.MODE_STITCHED, 0, B.WRAP_ZLIB), ("primed", B.MODE_INDEPENDENT, B.FLAG_PRIME, B.WRAP_RAW)):
        print(f"{mib} MiB {name} L{lvl} ...", flush=True)
        r = B.deflate_batch_dev(t, 65536, lvl, wrap, mode, flags)
        torch.cuda.synchronize()
        print("   ok", r.read_result().total_out_bytes, flush=True)
