"""GPU: the sharded entry point on one rank (world size 1) -- part flags, exchange plan, gather."""
import zlib

import numpy as np
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


def test_empty_middle_part(gpu_ctx, oracle):
    """A rank whose chunk range is empty but lies in the middle of the stream (fewer chunks than ranks):
    no input bytes, 32 KiB of history before them, neither first nor last -- found by the 8-rank run of
    tools/sharded_check.py (torch hands out a null data_ptr() for an empty view)."""
    import zlib, torch
    B = pkg("batch")
    S = pkg("sharded")
    data = torch.from_numpy(np.frombuffer(make_mixed(70000, 3), dtype=np.uint8).copy()).cuda()
    for wrap in (0, 1, 2):
        for level in (0, 1, 6):
            parts = []
            for r, (b0, b1) in enumerate(((0, 0), (0, 65536), (65536, 65536), (65536, 65536), (65536, 70000))):
                flags = (0 if r == 0 else B.FLAG_NOT_FIRST) | (0 if r == 4 else B.FLAG_NOT_LAST)
                res = B.deflate_batch_dev(data[b0:b1], 65536, level, wrap, B.MODE_STITCHED, flags, history=min(b0, 32768), want_checks=True)
                rr = res.read_result()
                parts.append((bytes(res.out[: (rr.total_out_bits + 7) // 8].cpu().numpy()), int(rr.total_out_bits), int(rr.check), b1 - b0))
            body, nbits = S.bit_concat_host([(p[0], p[1]) for p in parts])
            check, total = None, 0
            for _, _, c, ln in parts:
                if wrap == 1:
                    check = c if check is None else B.adler32_combine(check, c, ln)
                elif wrap == 2:
                    check = c if check is None else B.crc32_combine(check, c, ln)
                total += ln
            stream = body + S.wrapper_trailer(wrap, check or 0, total)
            d = zlib.decompressobj({0: -15, 1: 15, 2: 31}[wrap])
            host = bytes(data.cpu().numpy())
            assert d.decompress(stream) + d.flush() == host and d.eof, (wrap, level)


def test_host_part_pipelined_and_whole_stream(gpu_ctx, oracle):
    """zs_deflate_part / zs_deflate_batch in STITCHED mode from host buffers, large enough (>= 512 chunks) to
    go through the sliced H2D / kernels / D2H pipeline: slices are parts that end with the Z_SYNC_FLUSH
    marker, the history before a part is matched against, the trailer of a whole stream is written from
    the combined checksums."""
    import ctypes as C
    import torch
    B, S, capi = pkg("batch"), pkg("sharded"), pkg("capi")
    lib = capi.load()
    data = make_mixed(40 << 20, 9)
    # the whole stream through the host-buffer batch call, every wrapper
    for wrap, wb in ((0, -15), (1, 15), (2, 31)):
        r = B.deflate_batch(data, 65536, 6, wrap, B.MODE_STITCHED)
        d = zlib.decompressobj(wb)
        assert d.decompress(r.data) + d.flush() == data and d.eof and d.unused_data == b"", wrap
        assert r.total_out_bits == 8 * len(r.data)
        if wrap == 1:
            assert r.check == zlib.adler32(data)
        if wrap == 2:
            assert r.check == zlib.crc32(data)
        ref = len(zlib.compress(data, 6))
        assert len(r.data) <= 1.03 * ref
    # a middle part with 32 KiB of history, raw: decodes with that history as the preset dictionary
    cut = 3 * 65536 + 32768
    h = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    n = len(data) - cut
    cap = int(lib.zs_deflate_batch_bound(n, -(-n // 65536), 65536, 0, 1)) + 64
    out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    res = capi.DeflateResult()
    rc = lib.zs_deflate_part(gpu_ctx.handle, C.c_void_p(h.data_ptr() + cut), n, 32768, 65536, 6, 0,
                             B.FLAG_NOT_FIRST | B.FLAG_NOT_LAST, C.c_void_p(out.data_ptr()), cap, None, C.byref(res))
    assert rc == 0, gpu_ctx.last_error()
    part = bytes(out[: res.total_out_bytes].numpy())
    assert part[-4:] == b"\x00\x00\xff\xff" and res.total_out_bits == 8 * len(part)
    d = zlib.decompressobj(-15, zdict=data[cut - 32768: cut])
    assert d.decompress(part) == data[cut:] and not d.eof
    # the same part without its history is a little larger (the first 32 KiB find fewer matches)
    rc = lib.zs_deflate_part(gpu_ctx.handle, C.c_void_p(h.data_ptr() + cut), n, 0, 65536, 6, 0,
                             B.FLAG_NOT_FIRST | B.FLAG_NOT_LAST, C.c_void_p(out.data_ptr()), cap, None, C.byref(res))
    assert rc == 0 and res.total_out_bytes >= len(part)


def test_null_part_with_history_is_an_argument_error(gpu_ctx):
    """An empty part must still name its position: a null input pointer together with history is refused
    with Z_STREAM_ERROR instead of reaching the kernels (found by the 8-rank run of round 1)."""
    import torch
    B, capi = pkg("batch"), pkg("capi")
    lib = capi.load()
    out = torch.empty(4096, dtype=torch.uint8, device="cuda")
    res = torch.zeros(24, dtype=torch.uint8, device="cuda")
    rc = lib.zs_deflate_batch_dev(gpu_ctx.handle, None, 0, None, 1, 65536, 65536, 32768, 6, 1, B.MODE_STITCHED,
                                  B.FLAG_NOT_FIRST | B.FLAG_NOT_LAST, out.data_ptr(), out.numel(), None, None, None, res.data_ptr())
    assert rc == capi.Z_STREAM_ERROR and "history" in gpu_ctx.last_error()
    rc = lib.zs_deflate_batch_dev(gpu_ctx.handle, None, 5, None, 1, 65536, 65536, 0, 6, 1, B.MODE_STITCHED, 0,
                                  out.data_ptr(), out.numel(), None, None, None, res.data_ptr())
    assert rc == capi.Z_STREAM_ERROR
    # without history an empty first-and-last part is a valid (empty) stream
    rc = lib.zs_deflate_batch_dev(gpu_ctx.handle, None, 0, None, 1, 65536, 65536, 0, 6, 1, B.MODE_STITCHED, 0,
                                  out.data_ptr(), out.numel(), None, None, None, res.data_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    nbytes = int.from_bytes(bytes(res[:8].cpu().numpy()), "little")
    assert zlib.decompress(bytes(out[:nbytes].cpu().numpy())) == b""


def test_inflate_parts_round_trip(gpu_ctx, oracle):
    """The inverse split (sharded.inflate_part / inflate_sharded): the parts of one stream, as deflate_sharded
    leaves them on the ranks, are inflated one by one -- rank 0's with the wrapper's windowBits, the later ones
    raw with the 32 KiB before their range as preset dictionary -- and their checksums folded with *_combine.
    Ranks are emulated on one GPU: large parts (segment-parallel decoder, with and without flush points inside),
    small ones (one warp), an empty range in the middle."""
    import torch
    S, B, capi = pkg("sharded"), pkg("batch"), pkg("capi")
    data = make_mixed(6 << 20, 11)
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    chunk = 65536
    cuts_list = ([0, 40, 40, 90, 96], [0, 1, 2, 96], [0, 96])      # chunk ranges per emulated rank
    for wrap, kind, ref in ((1, capi.KIND_ADLER32, zlib.adler32(data)), (2, capi.KIND_CRC32, zlib.crc32(data)), (0, None, 0)):
        for cuts in cuts_list:
            for sync in (0, B.FLAG_SYNC):
                world = len(cuts) - 1
                check, total, pos = None, 0, 0
                for r in range(world):
                    b0, b1 = cuts[r] * chunk, cuts[r + 1] * chunk
                    hist = min(b0, 32768)
                    res = B.deflate_batch_dev(t[b0:b1], chunk, 6, wrap, B.MODE_STITCHED, S.part_flags(r, world) | sync, history=hist)
                    rr = res.read_result()
                    assert rr.total_out_bits == 8 * rr.total_out_bytes
                    ip = S.inflate_part(res.out, int(rr.total_out_bytes), b1 - b0, wrap, r == 0, t[b0 - hist: b0] if hist else None)
                    last = r == world - 1
                    # a lone part is a whole stream with its trailer; otherwise only the last part ends the stream
                    assert ip.status == (capi.Z_STREAM_END if last else capi.Z_BUF_ERROR), (wrap, cuts, r, ip.status)
                    assert ip.out_len == b1 - b0 and ip.in_used == rr.total_out_bytes, (wrap, cuts, r, ip.out_len, ip.in_used)
                    assert torch.equal(ip.out[: ip.out_len], t[b0:b1]), (wrap, cuts, r)
                    if kind is not None:
                        assert ip.check == int(rr.check)
                        check = ip.check if check is None else (B.adler32_combine if kind == capi.KIND_ADLER32 else B.crc32_combine)(check, ip.check, ip.out_len)
                    total += ip.out_len
                assert total == len(data)
                if kind is not None:
                    assert check == ref, (wrap, cuts)
    # the collective form on one rank (no process group): plan = the part's own numbers
    res, rr, plan = S.deflate_sharded(t, 262144, 6, 1)
    ip, iplan = S.inflate_sharded(res.out, int(rr.total_out_bytes), len(data), 1)
    assert ip.status == capi.Z_STREAM_END and torch.equal(ip.out[: ip.out_len], t)
    assert iplan.check == plan.check == zlib.adler32(data) and iplan.total_len == len(data) and iplan.total_bits == 8 * rr.total_out_bytes
