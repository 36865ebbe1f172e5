"""CPU suite, world_size 2 over gloo: the N>1 host path -- contiguous chunk ranges, the all_gather
of (bit length, checksum, length) per rank, the exclusive scan and the checksum fold.  The per-rank
parts are produced by the oracle (there is no GPU here), framed exactly like the GPU parts:
raw deflate of the rank's range primed with the preceding 32 KiB, non-final parts ending with the
Z_SYNC_FLUSH marker."""
import os
import socket
import zlib

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, make_text, pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, wrap, q, size=700000):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    S = pkg("sharded")
    capi = pkg("capi")
    data = make_text(size, 77)
    chunk = 65536
    n_chunks = -(-len(data) // chunk)
    lo, hi = S.shard_range(n_chunks, rank, world)
    b0, b1 = lo * chunk, min(hi * chunk, len(data))
    local = data[b0:b1]
    last = rank == world - 1
    part = O.deflate(local, 6, 0, data[max(0, b0 - 32768): b0] or None, O.Z_FINISH if last else O.Z_SYNC_FLUSH)
    if rank == 0:
        part = S.wrapper_header(wrap, 6) + part
    kind = None if wrap == 0 else (capi.KIND_ADLER32 if wrap == 1 else capi.KIND_CRC32)
    check = 0 if wrap == 0 else (zlib.adler32(local) if wrap == 1 else zlib.crc32(local))
    plan = S.exchange_meta(len(part) * 8, check, len(local), kind)
    gathered = [None] * world
    dist.all_gather_object(gathered, (part, len(part) * 8))
    ok = True
    if rank == 0:
        body, nbits = S.bit_concat_host(gathered)
        ok &= nbits == plan.total_bits and plan.bit_offset[0] == 0
        ok &= plan.bit_offset[1] == len(gathered[0][0]) * 8
        stream = body + S.wrapper_trailer(wrap, plan.check, plan.total_len)
        wb = {0: -15, 1: 15, 2: 31}[wrap]
        d = zlib.decompressobj(wb)
        ok &= d.decompress(stream) + d.flush() == data and d.eof
        ret, out, used, _ = O.inflate(stream, wb, len(data) + 64)
        ok &= ret == 1 and out == data and used == len(stream)
        if wrap == 1:
            ok &= plan.check == zlib.adler32(data)
        if wrap == 2:
            ok &= plan.check == zlib.crc32(data)
        ok &= plan.total_len == len(data)
    # The inverse split (sharded.inflate_sharded's contract, with the oracle standing in for the GPU decoder): every
    # rank inflates the part it holds -- rank 0 with the wrapper's windowBits, the others raw with the 32 KiB before
    # their range as preset dictionary -- only the last part ends with Z_STREAM_END, the others run out of input
    # behind their Z_SYNC_FLUSH marker (Z_BUF_ERROR under Z_FINISH) with all of their output written; the same
    # exchange step folds the checksums and lengths.
    wb = {0: -15, 1: 15, 2: 31}[wrap] if rank == 0 else -15
    ret, out, used, _ = O.inflate(part, wb, len(local) + 64, None if rank == 0 else (data[max(0, b0 - 32768): b0] or None))
    ok &= ret == (O.Z_STREAM_END if last else O.Z_BUF_ERROR) and out == local and used == len(part)
    out_check = 0 if wrap == 0 else (zlib.adler32(out) if wrap == 1 else zlib.crc32(out))
    iplan = S.exchange_meta(used * 8, out_check, len(out), kind)
    ok &= iplan.check == plan.check and iplan.total_len == len(data) and iplan.total_bits == plan.total_bits
    ok &= iplan.bit_offset == plan.bit_offset
    q.put((rank, bool(ok), plan.my_bit_offset))
    dist.barrier()
    dist.destroy_process_group()


def _run(wrap, world=2, size=700000):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, wrap, q, size)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_two_ranks_zlib_stream():
    res = _run(1)
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == 0 and res[1][2] > 0


def test_two_ranks_gzip_and_raw():
    assert all(ok for _, ok, _ in _run(2))
    assert all(ok for _, ok, _ in _run(0))


def test_three_ranks_with_an_empty_range():
    """Fewer chunks than ranks (2 chunks of 64 KiB over 3 ranks: rank 0's range is empty and it still writes the
    wrapper header, an empty block and the marker): the plan, the stitched stream and the inverse exchange need no
    special case."""
    res = _run(1, world=3, size=100000)
    assert len(res) == 3 and all(ok for _, ok, _ in res)
    assert res[0][2] == 0 and 0 < res[1][2] < res[2][2]


def test_shard_range_and_bit_concat():
    S = pkg("sharded")
    for n in (0, 1, 7, 16384):
        for w in (1, 2, 4, 8):
            cover = []
            for r in range(w):
                lo, hi = S.shard_range(n, r, w)
                cover += list(range(lo, hi))
            assert cover == list(range(n))
    # bit-granular concatenation at arbitrary (non byte) lengths
    a, b, c = (bytes([0b10110101, 0b1]), 9), (bytes([0xFF, 0x0F]), 12), (bytes([0b101]), 3)
    body, nbits = S.bit_concat_host([a, b, c])
    assert nbits == 24
    v = int.from_bytes(body, "little")
    assert v & 0x1FF == 0b110110101 and (v >> 9) & 0xFFF == 0xFFF and (v >> 21) == 0b101
    # single rank plan without a process group
    plan = S.exchange_meta(100, 5, 10, None, header_bits=16)
    assert plan.bit_offset == [16] and plan.total_bits == 116 and plan.my_bit_offset == 16
