"""GPU parity: K8 adler32 / crc32 (+ combine) against the oracle restatement of
src/mod/common/adler32.ts and crc32.ts -- bit-exact."""
import zlib

import numpy as np
import pytest

from conftest import make_text, pkg, rand_bytes

pytestmark = pytest.mark.gpu


def test_kats(gpu_ctx, oracle, kat):
    B = pkg("batch")
    for v in kat["crc32"]:
        assert B.checksum(bytes.fromhex(v["data_hex"]), 1, v["init"]) == v["expect"], v["ref"]
    for v in kat["adler32"]:
        assert B.checksum(bytes.fromhex(v["data_hex"]), 0, v["init"]) == v["expect"], v["ref"]
    assert B.checksum(b"", 0) == 1 and B.checksum(b"", 1) == 0


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 31, 32, 33, 511, 512, 513, 4096, 65535, 65536, 65537, 1000003, 9 << 20])
def test_whole_buffer(gpu_ctx, oracle, n):
    B = pkg("batch")
    d = rand_bytes(n, n) if n < (1 << 20) else make_text(n, n)
    assert B.checksum(d, 1) == oracle.crc32(d) == zlib.crc32(d)
    assert B.checksum(d, 0) == oracle.adler32(d) == zlib.adler32(d)
    # continuation from a passed value (the reference's running _adler)
    cut = n // 3
    assert B.checksum(d[cut:], 1, zlib.crc32(d[:cut])) == zlib.crc32(d)
    assert B.checksum(d[cut:], 0, zlib.adler32(d[:cut])) == zlib.adler32(d)
    # worst case for the adler sums
    ff = b"\xff" * n
    assert B.checksum(ff, 0) == zlib.adler32(ff)


def test_segments_ragged(gpu_ctx, oracle):
    import torch
    B = pkg("batch")
    rng = np.random.default_rng(5)
    lens = np.concatenate([[0, 1, 2, 3, 15, 16, 17, 4096, 70000], rng.integers(0, 9000, size=300)])
    off = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    data = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    t = torch.from_numpy(data).cuda()
    o = torch.from_numpy(off).cuda()
    crc = B.checksum_batch_dev(t, o, 1).cpu().numpy().view(np.uint32)
    adl = B.checksum_batch_dev(t, o, 0).cpu().numpy().view(np.uint32)
    for i in range(lens.size):
        seg = data[off[i]: off[i + 1]].tobytes()
        assert crc[i] == zlib.crc32(seg), (i, lens[i])
        assert adl[i] == zlib.adler32(seg), (i, lens[i])
    # checksum of checksums: folding the per-segment values equals the whole-buffer value
    acc_c, acc_a = 0, 1
    for i in range(lens.size):
        acc_c = B.crc32_combine(acc_c, int(crc[i]), int(lens[i]))
        acc_a = B.adler32_combine(acc_a, int(adl[i]), int(lens[i]))
    assert acc_c == zlib.crc32(data.tobytes()) and acc_a == zlib.adler32(data.tobytes())
    assert B.checksum_dev(t, 1) == acc_c and B.checksum_dev(t, 0) == acc_a


def test_long_segments_and_row_boundaries(gpu_ctx):
    """One warp per segment reads rows of 512 bytes: lengths around multiples of 16 and 512 at every start
    alignment, and single segments of tens of MiB (worst-case bytes for the adler32 accumulators)."""
    import torch
    B = pkg("batch")
    rng = np.random.default_rng(6)
    lens = []
    for base in (0, 16, 496, 512, 528, 1024, 5 * 512, 65536):
        for d in (-17, -16, -15, -1, 0, 1, 15, 16, 17):
            if base + d >= 0:
                lens.append(base + d)
    lens += [int(x) for x in rng.integers(0, 40, 40)]        # shifts the alignment of what follows
    lens += [20 << 20, (24 << 20) + 13]
    lens = np.array(lens, dtype=np.int64)
    rng.shuffle(lens[:-2])
    off = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    data = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    data[off[-3]: off[-2]] = 0xff                               # the 20 MiB segment: all 0xff
    t = torch.from_numpy(data).cuda()
    o = torch.from_numpy(off).cuda()
    crc = B.checksum_batch_dev(t, o, 1).cpu().numpy().view(np.uint32)
    adl = B.checksum_batch_dev(t, o, 0).cpu().numpy().view(np.uint32)
    for i in range(lens.size):
        seg = data[off[i]: off[i + 1]].tobytes()
        assert crc[i] == zlib.crc32(seg), (i, lens[i], off[i] % 16)
        assert adl[i] == zlib.adler32(seg), (i, lens[i], off[i] % 16)
