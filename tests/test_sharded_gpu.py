"""GPU: the sharded entry point on one rank (world size 1) -- part flags, exchange plan, gather."""
import zlib

import pytest

from conftest import make_mixed, pkg

pytestmark = pytest.mark.gpu


def test_single_rank_plan_and_gather(gpu_ctx, oracle):
    import torch
    S, B = pkg("sharded"), pkg("batch")
    data = make_mixed(3 << 20, 5)
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    for wrap in (0, 1, 2):
        res, rr, plan = S.deflate_sharded(t, 262144, 6, wrap)
        assert plan.bit_offset == [0] and plan.total_bits == rr.total_out_bits and plan.total_len == len(data)
        stream = res.out[: rr.total_out_bytes].cpu().numpy().tobytes()
        wb = {0: -15, 1: 15, 2: 31}[wrap]
        d = zlib.decompressobj(wb)
        assert d.decompress(stream) + d.flush() == data and d.eof
        if wrap == 1:
            assert plan.check == zlib.adler32(data)
        if wrap == 2:
            assert plan.check == zlib.crc32(data)


def test_two_parts_in_one_process(gpu_ctx, oracle):
    # the two ranks of a 2-GPU run, emulated on one GPU: part flags + history + bit-granular stitch
    import ctypes as C
    import torch
    S, B, capi = pkg("sharded"), pkg("batch"), pkg("capi")
    data = make_mixed(2 << 20, 6)
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    chunk = 65536
    n_chunks = len(data) // chunk
    parts, metas = [], []
    for r in range(2):
        lo, hi = S.shard_range(n_chunks, r, 2)
        res = B.deflate_batch_dev(t[lo * chunk: hi * chunk], chunk, 6, B.WRAP_ZLIB, B.MODE_STITCHED, S.part_flags(r, 2),
                                  history=min(lo * chunk, 32768))
        rr = res.read_result()
        parts.append((res.out[: rr.total_out_bytes].cpu().numpy().tobytes(), int(rr.total_out_bits)))
        metas.append((int(rr.check), (hi - lo) * chunk))
    body, nbits = S.bit_concat_host(parts)
    check = capi.load().zs_adler32_combine(metas[0][0], metas[1][0], metas[1][1])
    stream = body + S.wrapper_trailer(1, check, len(data))
    assert zlib.decompress(stream) == data
    assert stream[:2] == S.wrapper_header(1, 6)
