"""torchrun --nproc-per-node N tools/sharded_check.py : N-rank stitched deflate of one replicated
input, gathered on rank 0 and decoded with C zlib + the oracle, and inflated again where it lies (every rank the
part it holds, sharded.inflate_sharded; added after the last multi-GPU run of round 2: the 2-rank bench line and
the one-GPU emulation in tests/test_sharded_gpu.py are its evidence so far).  Run on the GPU box."""
import importlib, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S = importlib.import_module("zlib-streams-ts_b200.sharded")
B = importlib.import_module("zlib-streams-ts_b200.batch")
corpus = importlib.import_module("zlib-streams-ts_b200.corpus")
ok = True


def inverse_ok(data, res, rr, plan, chunk, wrap):
    """The inverse split, sharded.inflate_sharded spelled out: every rank inflates the part it holds, the ranks fold
    checksums and lengths; all ranks get the verdict."""
    n = data.numel()
    lo, hi = S.shard_range(B.n_chunks_for(n, chunk), rank, world)
    b0, b1 = lo * chunk, min(hi * chunk, n)
    hist = min(b0, 32768)
    capi = importlib.import_module("zlib-streams-ts_b200.capi")
    kind = None if wrap == 0 else (capi.KIND_ADLER32 if wrap == 1 else capi.KIND_CRC32)
    ip = None
    try:   # the GPU call alone may fail on one rank: the exchange below runs on every rank whatever happened
        ip = S.inflate_part(res.out, int(rr.total_out_bytes), b1 - b0, wrap, rank == 0, data[b0 - hist: b0] if hist else None)
    except Exception as e:
        print("inflate_part raised on rank", rank, repr(e))
    meta = (ip.in_used * 8, ip.check, ip.out_len) if ip is not None else (0, 0, 0)
    iplan = S.exchange_meta(meta[0], meta[1], meta[2], kind, 0, device=data.device)
    want = 1 if rank == world - 1 else -5
    good = (ip is not None and ip.status == want and ip.out_len == b1 - b0 and ip.in_used == int(rr.total_out_bytes)
            and bool(torch.equal(ip.out[: ip.out_len], data[b0:b1])) and iplan.total_len == n
            and (wrap == 0 or iplan.check == plan.check))
    t = torch.tensor([1 if good else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


for wrap, level, chunk, n in ((1, 6, 262144, 24 << 20), (2, 1, 65536, 10 << 20), (0, 6, 65536, (5 << 20) + 12345)):
    data = torch.from_numpy(corpus.mixed_numpy(n, 0xB200)).cuda()     # same bytes on every rank
    res, rr, plan = S.deflate_sharded(data, chunk, level, wrap)
    stream = S.gather_stream(res, rr, plan, wrap, dst=0)
    inv = inverse_ok(data, res, rr, plan, chunk, wrap)
    if rank == 0:
        host = data.cpu().numpy().tobytes()
        wb = {0: -15, 1: 15, 2: 31}[wrap]
        d = zlib.decompressobj(wb)
        good = d.decompress(stream) + d.flush() == host and d.eof
        if wrap == 1: good &= plan.check == zlib.adler32(host)
        if wrap == 2: good &= plan.check == zlib.crc32(host)
        ref = len(zlib.compress(host, level))
        print(f"wrap {wrap} level {level} world {world}: ok={good} inflate_sharded={inv} size {len(stream)} vs zlib {ref} ({len(stream)/ref:.4f}) bit offsets {plan.bit_offset}")
        ok &= good and inv
    dist.barrier()
# randomised part (argument: seconds): sizes down to fewer chunks than ranks, every level, both chunk sizes
import numpy as np, time
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
rng = np.random.default_rng(11)     # same sequence on every rank
t0 = time.time(); it = 0
while True:
    go = torch.tensor([1 if time.time() - t0 < budget else 0], device="cuda")
    dist.broadcast(go, 0)
    if not int(go.item()):
        break
    it += 1
    n = int(rng.choice([1, 100, 70000, 300000, 2 << 20, (9 << 20) + 77]))
    wrap, level = int(rng.integers(0, 3)), int(rng.integers(0, 10))
    chunk = int(rng.choice([4096, 65536, 262144]))
    seed = int(rng.integers(1 << 30))
    data = torch.from_numpy(corpus.mixed_numpy(n, seed)).cuda()
    res, rr, plan = S.deflate_sharded(data, chunk, level, wrap)
    stream = S.gather_stream(res, rr, plan, wrap, dst=0)
    inv = inverse_ok(data, res, rr, plan, chunk, wrap)
    if rank == 0:
        host = data.cpu().numpy().tobytes()
        d = zlib.decompressobj({0: -15, 1: 15, 2: 31}[wrap])
        good = d.decompress(stream) + d.flush() == host and d.eof and inv
        if not good:
            print("random case FAILED", it, n, wrap, level, chunk, seed)
        ok &= good
    dist.barrier()
if rank == 0:
    print("random cases:", it)
    print("SHARDED_CHECK", "PASS" if ok else "FAIL")
dist.destroy_process_group()
