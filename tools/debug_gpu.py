import sys, os, zlib, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
print("== checksum alignment/size sweep")
rng = np.random.default_rng(1)
data = rng.integers(0, 256, size=400000, dtype=np.uint8)
t = torch.from_numpy(data).cuda()
for ln in (4096, 32768, 65535, 65536, 65537, 70000, 131072, 200000):
    for beg in (0, 6, 16, 54):
        off = torch.tensor([beg, beg + ln], dtype=torch.int64, device="cuda")
        c = int(B.checksum_batch_dev(t, off, 1).cpu().numpy().view(np.uint32)[0])
        a = int(B.checksum_batch_dev(t, off, 0).cpu().numpy().view(np.uint32)[0])
        seg = data[beg:beg + ln].tobytes()
        print(ln, beg, "crc", "ok" if c == zlib.crc32(seg) else f"BAD {c:08x} {zlib.crc32(seg):08x}", "adler", "ok" if a == zlib.adler32(seg) else f"BAD {a:08x} {zlib.adler32(seg):08x}")
print("== stitched deflate sweep")
for n, chunk in ((4 << 20, 65536), (10 << 20, 65536), (12 << 20, 262144), (40 << 20, 262144), (48 << 20, 262144)):
    tt = corpus.text_torch(n, torch.device("cuda:0"), seed=77)
    host = tt.cpu().numpy().tobytes()
    for lvl in (1, 6):
        res = B.deflate_batch_dev(tt, chunk, lvl, B.WRAP_ZLIB, B.MODE_STITCHED)
        rr = res.read_result()
        stream = res.out[: rr.total_out_bytes].cpu().numpy().tobytes()
        d = zlib.decompressobj()
        try:
            out = d.decompress(stream)
            ok = out == host and d.eof
            msg = f"eof={d.eof} outlen={len(out)} match_prefix={out == host[:len(out)]} unused={len(d.unused_data)}"
        except zlib.error as e:
            ok = False
            # find how far it decodes
            d = zlib.decompressobj(); got = b""
            try:
                for i in range(0, len(stream), 4096):
                    got += d.decompress(stream[i:i + 4096])
            except zlib.error:
                pass
            msg = f"{e}; decoded {len(got)} bytes, prefix ok={got == host[:len(got)]} (chunk {len(got) // chunk})"
        nblk = rr.n_blocks
        ob = res.out_bits.cpu().numpy(); oo = res.out_off.cpu().numpy()
        print(n, chunk, lvl, "OK" if ok else "FAIL", msg, "bytes", rr.total_out_bytes, "bits", rr.total_out_bits, "blocks", nblk, "sum_bits", int(ob.sum()), "last_off", int(oo[-1]))
