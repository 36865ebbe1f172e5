"""CPU: the bit tricks of par_spec_kernel (csrc/zs_inflate_par.cu), restated in Python and checked against the plain
definition of what they test -- the first 17 bits of a dynamic block header (inflate.ts:684-690: BFINAL = 0, BTYPE = 2,
HLIT <= 29, HDIST <= 29) for 32 consecutive bit positions at once, and the Kraft sum of the code-length code from a
byte table held in a 64-bit constant.  The constant is read from the CUDA source, so the two cannot drift apart."""
import os
import random
import re

from conftest import ROOT

SRC = open(os.path.join(ROOT, "zlib-streams-ts_b200", "csrc", "zs_inflate_par.cu")).read()
M64 = (1 << 64) - 1


def quick_mask(w):   # spec_quick_mask
    m = ~w & ~(w >> 1) & (w >> 2)
    m &= ~((w >> 4) & (w >> 5) & (w >> 6) & (w >> 7))
    m &= ~((w >> 9) & (w >> 10) & (w >> 11) & (w >> 12))
    return m & 0xFFFFFFFF


def test_the_source_still_uses_these_formulas():
    body = SRC[SRC.index("spec_quick_mask(uint64_t w)"):]
    body = body[: body.index("}")]
    assert "~w & ~(w >> 1) & (w >> 2)" in body
    assert "(w >> 4) & (w >> 5) & (w >> 6) & (w >> 7)" in body and "(w >> 9) & (w >> 10) & (w >> 11) & (w >> 12)" in body


def test_quick_mask_is_the_17_bit_header_test():
    rng = random.Random(11)
    for _ in range(5000):
        w = rng.getrandbits(64)
        m = quick_mask(w)
        for i in range(32):
            v = w >> i
            want = (v & 7) == 4 and ((v >> 3) & 31) <= 29 and ((v >> 8) & 31) <= 29
            assert ((m >> i) & 1) == int(want)
    # at most 11 finds per 32 positions (the pattern 0,0,1 is three bits long): the bound of the kernel's queue
    assert max(bin(quick_mask(rng.getrandbits(64))).count("1") for _ in range(20000)) <= 11
    assert "kSpecQueue = 32 + 32 * 11" in SRC


def test_kraft_table_constant():
    weights = int(re.search(r"const uint64_t weights = (0x[0-9a-fA-F]+)ull", SRC).group(1), 16)
    assert [(weights >> (8 * l)) & 0xFF for l in range(8)] == [0] + [128 >> l for l in range(1, 8)]
    rng = random.Random(12)
    for _ in range(5000):
        c, ncode = rng.getrandbits(57), rng.randint(4, 19)
        masked = c & ((1 << (3 * ncode)) - 1) if ncode < 19 else c
        fast = sum((weights >> (8 * ((masked >> (3 * i)) & 7))) & 0xFF for i in range(19))
        plain = sum((128 >> l) for l in (((c >> (3 * i)) & 7) for i in range(ncode)) if l)
        assert fast == plain
