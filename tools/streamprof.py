# The drop-in stream API (zs_stream_deflate / zs_stream_inflate, 32 KiB in / 64 KiB out like streams.ts) with the
# per-kernel times of the calls behind it: where a DecompressionStream's time goes.
import sys, os, importlib, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
B = importlib.import_module("zlib-streams-ts_b200.batch")
capi = importlib.import_module("zlib-streams-ts_b200.capi")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = B.default_context(0)
host = corpus.text_numpy(mib << 20, 5)
ctx.profile(True); ctx.profile_read()
t0 = time.perf_counter()
out = bench.stream_api_bench(ctx, host, capi)
dt = time.perf_counter() - t0
prof = ctx.profile_read(); ctx.profile(False)
print(json.dumps(out, indent=1))
tot = sum(v[1] for v in prof.values())
print(f"wall {dt:.3f} s, kernel time {tot/1e3:.3f} s")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {k:32s} {v[0]:6d} launches {v[1]:10.2f} ms")
