"""sha256 of the deflate output over a set of shapes (regression check between two builds)."""
import hashlib, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for kind in ("text", "mixed"):
    data = (corpus.text_torch(mib << 20, "cuda") if kind == "text" else corpus.mixed_torch(mib << 20, "cuda"))
    for (chunk, level, wrap, mode, flags) in [(4096, 6, B.WRAP_GZIP, B.MODE_INDEPENDENT, 0), (65536, 1, B.WRAP_RAW, B.MODE_INDEPENDENT, B.FLAG_PRIME),
                                              (65536, 6, B.WRAP_ZLIB, B.MODE_STITCHED, 0), (262144, 6, B.WRAP_ZLIB, B.MODE_STITCHED, 0),
                                              (100000, 4, B.WRAP_RAW, B.MODE_INDEPENDENT, 0), (1000, 1, B.WRAP_RAW, B.MODE_INDEPENDENT, B.FLAG_PRIME)]:
        r = B.deflate_batch_dev(data, chunk_size=chunk, level=level, wrap=wrap, mode=mode, flags=flags)
        torch.cuda.synchronize()
        total = int(r.out_off[-1].item()) if mode == B.MODE_INDEPENDENT else (int(r.out_off[-1].item()) + 7) // 8 + 8
        out = r.out[:total].cpu().numpy().tobytes()
        print(kind, chunk, level, wrap, mode, flags, total, hashlib.sha256(out).hexdigest()[:16], hashlib.sha256(r.out_bits.cpu().numpy().tobytes()).hexdigest()[:16])
