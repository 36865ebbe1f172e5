# driver for ncu captures of the one-stream decoder (zs_inflate_par.cu): N MiB of text as ONE C zlib stream written in
# one go (no flush points: par_spec_kernel locates the block headers), decoded once by zs_inflate_stream_dev.
#   ncu --set full -k regex:par_ -c 12 --import-source on -o gpurun_out/par_r2c python tools/ncu_par.py 128
import ctypes as C, importlib, os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
capi = importlib.import_module("zlib-streams-ts_b200.capi")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
lib = capi.load()
dev = torch.device("cuda:0")
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 128) << 20
host = corpus.text_numpy(n, 3).tobytes()
z = zlib.compress(host, 6)
ctx = B.default_context(0)
d_in = torch.frombuffer(bytearray(z + bytes(8)), dtype=torch.uint8).to(dev)
d_out = torch.empty(n + 64, dtype=torch.uint8, device=dev)
res = torch.zeros(4, dtype=torch.int64, device=dev)
rc = lib.zs_inflate_stream_dev(ctx.handle, d_in.data_ptr(), len(z), 15, d_out.data_ptr(), d_out.numel(), res.data_ptr(),
                               res.data_ptr() + 8, res.data_ptr() + 16, res.data_ptr() + 24, None, 0)
ctx.check(rc, "zs_inflate_stream_dev")
torch.cuda.synchronize()
out_len, in_used, check, status = (int(x) for x in res.cpu())
print("ok", status & 0xffffffff == 1 and out_len == n and bytes(d_out[:n].cpu().numpy()) == host, len(z))
