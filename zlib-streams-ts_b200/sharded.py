"""Contiguous chunk ranges over the ranks of one box (one process per GPU, torch.distributed).

The reference is single-threaded; sharding is new.  Rank r deflates chunks
[r*N/R, (r+1)*N/R) of the input as one *part* of a single stream (ZS_MODE_STITCHED with
ZS_FLAG_NOT_FIRST / ZS_FLAG_NOT_LAST, primed with the 32 KiB that precede its range), and the
ranks then run the path's only exchange step: an all_gather of (compressed bit length, checksum,
uncompressed length) per rank -> exclusive scan of the bit lengths (each rank's global bit offset)
and crc32_combine / adler32_combine of the checksums.  The payload stays on the GPU that produced
it unless one contiguous buffer is asked for (gather_stream: byte gather + bit-shift stitch with
zs_bit_concat_dev on the destination rank).

The collective works on any backend: NCCL on the GPUs, gloo in the CPU tests (tests/test_sharded.py).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import capi


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition: rank r gets [r*N/R, (r+1)*N/R)."""
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


@dataclass
class StitchPlan:
    bit_offset: list          # global bit offset of every rank's part
    total_bits: int
    my_bit_offset: int
    check: int                # checksum of the whole input (adler32 or crc32)
    total_len: int


def exchange_meta(part_bits: int, check: int, in_len: int, kind: int | None, header_bits: int = 0,
                  device=None, group=None) -> StitchPlan:
    """all_gather of (bits, check, len) per rank, exclusive scan, checksum fold.

    `kind`: capi.KIND_ADLER32, capi.KIND_CRC32 or None (raw stream, no checksum).
    Without an initialised process group this is the single-rank identity.
    """
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        mine = torch.tensor([part_bits, check, in_len], dtype=torch.int64, device=device)
        allv = torch.empty(world * 3, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allv, mine, group=group)
        rows = allv.view(world, 3).cpu().tolist()
    else:
        world, rank = 1, 0
        rows = [[part_bits, check, in_len]]
    offs, pos = [], header_bits
    for bits, _, _ in rows:
        offs.append(pos)
        pos += bits
    lib = capi.load()
    acc = None
    total_len = 0
    for _, ck, ln in rows:
        if kind is not None:
            if acc is None:
                acc = ck & 0xFFFFFFFF
            elif kind == capi.KIND_CRC32:
                acc = lib.zs_crc32_combine(acc, ck & 0xFFFFFFFF, ln)
            else:
                acc = lib.zs_adler32_combine(acc, ck & 0xFFFFFFFF, ln)
        total_len += ln
    return StitchPlan(offs, pos, offs[rank], acc if acc is not None else 0, total_len)


def part_flags(rank: int, world: int) -> int:
    f = 0
    if rank > 0:
        f |= capi.FLAG_NOT_FIRST
    if rank < world - 1:
        f |= capi.FLAG_NOT_LAST
    return f


def wrapper_header(wrap: int, level: int) -> bytes:
    """Header bytes of the stitched stream (deflate.ts:750-800); written by rank 0's part."""
    if wrap == capi.WRAP_ZLIB:
        h = (8 + (7 << 4)) << 8
        h |= (0 if level < 2 else 1 if level < 6 else 2 if level == 6 else 3) << 6
        h += 31 - h % 31
        return bytes([h >> 8, h & 0xFF])
    if wrap == capi.WRAP_GZIP:
        return bytes([0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 2 if level == 9 else 4 if level < 2 else 0, 255])
    return b""


def wrapper_trailer(wrap: int, check: int, total_len: int) -> bytes:
    """Trailer of the stitched stream (deflate.ts:964-988) from the folded checksum."""
    if wrap == capi.WRAP_ZLIB:
        return (check & 0xFFFFFFFF).to_bytes(4, "big")
    if wrap == capi.WRAP_GZIP:
        return (check & 0xFFFFFFFF).to_bytes(4, "little") + (total_len & 0xFFFFFFFF).to_bytes(4, "little")
    return b""


def bit_concat_host(parts) -> tuple[bytes, int]:
    """Reference implementation of the stitch for tests: parts = [(bytes, n_bits), ...]."""
    acc, nbits = 0, 0
    for data, bits in parts:
        v = int.from_bytes(data[: (bits + 7) // 8], "little") & ((1 << bits) - 1)
        acc |= v << nbits
        nbits += bits
    return acc.to_bytes((nbits + 7) // 8, "little"), nbits


def deflate_sharded(d_all_or_local: torch.Tensor, chunk_size: int, level: int, wrap: int, local_is_shard: bool = False,
                    history: int = 0, group=None, ctx=None, reuse=None):
    """Deflate this rank's contiguous chunk range of a stream and run the exchange step (GPU).

    If `local_is_shard` the tensor is already this rank's range (with `history` valid bytes before
    it in the same storage); otherwise it is the whole input, replicated, and the rank slices it.
    Returns (DeflateBatchDev, DeflateResult, StitchPlan).  The part starts at bit 0 of its own
    buffer; rank 0's part includes the wrapper header; the trailer comes from wrapper_trailer().
    `reuse` = the DeflateBatchDev of an earlier call with the same shapes (no allocation).
    """
    from . import batch as B
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if local_is_shard:
        local, hist = d_all_or_local, history
    else:
        n = d_all_or_local.numel()
        n_chunks = B.n_chunks_for(n, chunk_size)
        lo, hi = shard_range(n_chunks, rank, world)
        b0, b1 = lo * chunk_size, min(hi * chunk_size, n)
        local, hist = d_all_or_local[b0:b1], min(b0, 32768)
    flags = part_flags(rank, world)
    res = B.deflate_batch_dev(local, chunk_size, level, wrap, B.MODE_STITCHED, flags, history=hist, ctx=ctx, reuse=reuse)
    rr = res.read_result()
    kind = None if wrap == capi.WRAP_RAW else (capi.KIND_ADLER32 if wrap == capi.WRAP_ZLIB else capi.KIND_CRC32)
    plan = exchange_meta(int(rr.total_out_bits), int(rr.check), local.numel(), kind, 0, device=local.device, group=group)
    return res, rr, plan


def deflate_sharded_host(h_buf: torch.Tensor, history: int, chunk_size: int, level: int, wrap: int, h_out: torch.Tensor,
                         device=None, group=None, ctx=None):
    """The same step from HOST buffers (what a Node-API caller on each rank does): h_buf holds `history`
    bytes of the stream followed by this rank's range; the part is deflated through zs_deflate_part (H2D,
    kernels and D2H pipelined inside the call) into h_out, then the ranks exchange (bits, check, length).
    Returns (DeflateResult, StitchPlan); the part's bytes are h_out[:result.total_out_bytes]."""
    import ctypes as C

    from . import batch as B
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ctx = ctx or B.default_context(device.index if device is not None else None)
    n = h_buf.numel() - history
    res = capi.DeflateResult()
    rc = capi.load().zs_deflate_part(ctx.handle, C.c_void_p(h_buf.data_ptr() + history), n, history, chunk_size, level, wrap,
                                     part_flags(rank, world), C.c_void_p(h_out.data_ptr()), h_out.numel(), None, C.byref(res))
    ctx.check(rc, "zs_deflate_part")
    kind = None if wrap == capi.WRAP_RAW else (capi.KIND_ADLER32 if wrap == capi.WRAP_ZLIB else capi.KIND_CRC32)
    plan = exchange_meta(int(res.total_out_bits), int(res.check), n, kind, 0, device=device, group=group)
    return res, plan


@dataclass
class InflatePart:
    out: torch.Tensor         # uint8: this rank's range of the stream's output (capacity >= out_len)
    result: torch.Tensor      # int64 [4] on the device: out_len, in_used, check (unused), status
    out_len: int
    in_used: int
    status: int               # Z_STREAM_END on the rank that holds the final block, Z_BUF_ERROR (input ran dry) before it
    check: int                # adler32 / crc32 of this rank's output (0 for raw streams)


def inflate_part(part: torch.Tensor, part_bytes: int, out_cap: int, wrap: int, first: bool,
                 d_hist: torch.Tensor | None = None, ctx=None, reuse: InflatePart | None = None) -> InflatePart:
    """Inflate ONE part of a sharded stream on this GPU (no communication): see inflate_sharded.  `first`: the
    part begins with the wrapper header (rank 0's part)."""
    import ctypes as C

    from . import batch as B
    assert part.is_cuda and part.dtype == torch.uint8 and part.data_ptr() % 8 == 0
    dev = part.device
    ctx = ctx or B.default_context(dev.index)
    if reuse is not None:
        out, result = reuse.out, reuse.result
    else:
        out = torch.empty(out_cap + 64, dtype=torch.uint8, device=dev)
        result = torch.zeros(4, dtype=torch.int64, device=dev)
    window_bits = -15
    if first:
        window_bits = -15 if wrap == capi.WRAP_RAW else 15 if wrap == capi.WRAP_ZLIB else 31
        d_hist = None
    p = result.data_ptr()
    hist_len = d_hist.numel() if d_hist is not None else 0
    rc = capi.load().zs_inflate_stream_dev(ctx.handle, C.c_void_p(part.data_ptr()), part_bytes, window_bits,
                                           C.c_void_p(out.data_ptr()), out_cap, C.c_void_p(p), C.c_void_p(p + 8),
                                           C.c_void_p(p + 16), C.c_void_p(p + 24),
                                           C.c_void_p(d_hist.data_ptr()) if hist_len else None, hist_len)
    ctx.check(rc, "zs_inflate_stream_dev")
    out_len, in_used, _, status = (int(x) for x in result.cpu())
    status &= 0xFFFFFFFF
    if status & 0x80000000:
        status -= 1 << 32
    kind = None if wrap == capi.WRAP_RAW else (capi.KIND_ADLER32 if wrap == capi.WRAP_ZLIB else capi.KIND_CRC32)
    check = 0
    if kind is not None:   # (an empty range: the checksum of nothing is its initial value)
        check = B.checksum_dev(out[:out_len], kind, ctx=ctx) if out_len else (1 if kind == capi.KIND_ADLER32 else 0)
    return InflatePart(out, result, out_len, in_used, status, check)


def inflate_sharded(part: torch.Tensor, part_bytes: int, out_cap: int, wrap: int, d_hist: torch.Tensor | None = None,
                    group=None, ctx=None, reuse: InflatePart | None = None):
    """The inverse of deflate_sharded: every rank inflates the part of the stream it holds (GPU), then the ranks
    fold their checksums and lengths with the same exchange step.

    `part` is what deflate_sharded left on this rank (bit 0 of the part = byte 0 of the tensor: parts of one
    stream are byte aligned, every non-last part ends with the Z_SYNC_FLUSH marker of deflate.ts:945-946), 8-byte
    aligned, `part_bytes` long.  Rank 0's part begins with the wrapper header and is decoded with the wrapper's
    windowBits; every later part is raw deflate whose back references reach into the previous rank's output, so
    it is given `d_hist`, the <= 32 KiB of the stream that precede its range (what deflate_sharded primed it with),
    as a preset dictionary (inflateSetDictionary, inflate.ts:1220).  A part is ONE stream for the decoder: it goes
    through the segment-parallel path of zs_inflate_stream_dev (cut at flush points, else at located block
    headers).  Only the last rank's part holds the final block, so it alone ends with Z_STREAM_END; the others run
    out of input at their marker (Z_BUF_ERROR under Z_FINISH, inflate.ts:1092-1098) with all of their output written.
    The wrapper trailer is not part of any part (wrapper_trailer() makes it from the folded checksum): the caller
    compares plan.check with the trailer it holds.

    Returns (InflatePart, StitchPlan): plan.check = checksum of the whole output, plan.total_len its length,
    plan.bit_offset[r] = 8 x the compressed bytes the ranks before r consumed.
    """
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ip = inflate_part(part, part_bytes, out_cap, wrap, rank == 0, d_hist, ctx=ctx, reuse=reuse)
    kind = None if wrap == capi.WRAP_RAW else (capi.KIND_ADLER32 if wrap == capi.WRAP_ZLIB else capi.KIND_CRC32)
    plan = exchange_meta(ip.in_used * 8, ip.check, ip.out_len, kind, 0, device=part.device, group=group)
    return ip, plan


def gather_stream(res, rr, plan: StitchPlan, wrap: int, dst: int = 0, group=None):
    """One contiguous stream on rank `dst`: byte gather of the parts + bit-shift stitch (K9)."""
    import ctypes as C

    from . import batch as B
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    my_bytes = (int(rr.total_out_bits) + 7) // 8
    sizes = [(b - a + 7) // 8 for a, b in zip(plan.bit_offset, plan.bit_offset[1:] + [plan.total_bits])]
    dev = res.out.device
    if rank != dst:
        dist.send(res.out[:my_bytes].contiguous(), dst, group=group)
        return None
    total_bytes = (plan.total_bits + 7) // 8
    out = torch.zeros(total_bytes + 16, dtype=torch.uint8, device=dev)
    ctx = B.default_context(dev.index)
    lib = capi.load()
    for r in range(world):
        if r == dst:
            part = res.out[:my_bytes]
        else:
            part = torch.empty(sizes[r] + 16, dtype=torch.uint8, device=dev)
            dist.recv(part[: sizes[r]], r, group=group)
        nbits = (plan.bit_offset[r + 1] if r + 1 < world else plan.total_bits) - plan.bit_offset[r]
        ctx.check(lib.zs_bit_concat_dev(ctx.handle, C.c_void_p(out.data_ptr()), plan.bit_offset[r],
                                        C.c_void_p(part.data_ptr()), nbits), "zs_bit_concat_dev")
    torch.cuda.synchronize(dev)
    body = out[:total_bytes].cpu().numpy().tobytes()
    if world == 1:
        return body   # a lone part is a whole stream: the engine framed it, trailer included
    return body + wrapper_trailer(wrap, plan.check, plan.total_len)
