# checksum throughput (device-resident) vs the measured HBM copy bandwidth
import sys, os, importlib, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
dev = torch.device("cuda:0")
peak = 6551.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for n in (64 << 20, 1 << 30, 4 << 30):
    t = torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev)
    for seg in (4096, 65536, 1 << 20):
        off = torch.arange(0, n + 1, seg, dtype=torch.int64, device=dev)
        for kind, name in ((0, "adler32"), (1, "crc32")):
            B.checksum_batch_dev(t, off, kind)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                B.checksum_batch_dev(t, off, kind)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{n >> 20:5d} MiB, {seg:7d}-byte segments, {name:7s}: {n / ms / 1e6:8.1f} GB/s = {n / ms / 1e6 / peak * 100:5.1f} % of {peak:.0f}", flush=True)
    del t
