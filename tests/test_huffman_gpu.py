"""Parity of the Huffman stage (rows a8, a9, a11 of the scope table): code lengths and canonical
codes of the literal/length and distance trees must equal the reference's build_tree / gen_bitlen /
gen_codes (src/mod/deflate/trees.ts:54-76,167-316) for the same symbol frequencies -- including the
heap tie-breaks and the overflow repair of length-limited codes."""
import ctypes as C

import numpy as np
import pytest

from conftest import pkg

pytestmark = pytest.mark.gpu


def _oracle_tree(oracle, kind, freq):
    n = 286 if kind == 0 else 30
    f = np.zeros(n, np.uint16)
    f[:] = freq[:n]
    lens = np.zeros(n, np.uint16)
    codes = np.zeros(n, np.uint16)
    ol, sl = C.c_uint32(0), C.c_uint32(0)
    oracle.lib().zo_build_tree(kind, f.ctypes.data, lens.ctypes.data, codes.ctypes.data, C.byref(ol), C.byref(sl))
    return lens, codes


def _cases():
    rnd = np.random.default_rng(11)
    cases = []
    for k in range(48):
        f = np.zeros(320, np.uint32)
        shape = k % 8
        if shape == 0:      # flat small counts: many heap ties
            f[:286] = rnd.integers(0, 4, 286); f[288:318] = rnd.integers(0, 3, 30)
        elif shape == 1:    # text-like: few symbols, skewed
            idx = rnd.choice(np.arange(97, 123), 20, replace=False)
            f[idx] = rnd.integers(1, 2000, 20); f[32] = 3000; f[257:270] = rnd.integers(0, 300, 13); f[288:310] = rnd.integers(0, 200, 22)
        elif shape == 2:    # geometric: forces the 15-bit overflow repair
            v = (1.7 ** np.arange(30)).astype(np.int64)
            f[rnd.choice(256, 30, replace=False)] = np.minimum(v, 16000)
            f[288:300] = np.minimum((2 ** np.arange(12)), 4000)
        elif shape == 3:    # Fibonacci-like literals (deepest possible tree)
            a, b, fib = 1, 1, []
            for _ in range(22):
                fib.append(a); a, b = b, a + b
            f[:22] = np.minimum(np.array(fib), 16000); f[288] = 5
        elif shape == 4:    # all 256 literals, uniform (incompressible block)
            f[:256] = rnd.integers(50, 70, 256)
        elif shape == 5:    # a single literal, no distances (the >= 2 codes rule)
            f[rnd.integers(0, 256)] = rnd.integers(1, 9000)
        elif shape == 6:    # one match symbol only
            f[257 + rnd.integers(0, 29)] = 40; f[288 + rnd.integers(0, 30)] = 40
        else:               # everything used
            f[:286] = rnd.integers(1, 60, 286); f[288:318] = rnd.integers(1, 500, 30)
        f[256] = 0          # END_BLOCK is counted by the engine
        cases.append(f)
    return np.stack(cases)


@pytest.mark.parametrize("kernel", ["warp", "thread"])
def test_code_lengths_and_codes_equal_reference(gpu_ctx, oracle, monkeypatch, kernel):
    # both Huffman kernels (warp per block / thread per block; the engine picks by block count)
    monkeypatch.setenv("ZS_HUFF_TPB" if kernel == "thread" else "ZS_HUFF_WARP", "1")
    B = pkg("batch")
    freq = _cases()
    in_len = np.full(freq.shape[0], 1 << 30, np.uint32) & 0xffff0000   # never "stored"
    code, typ, bits = B.huffman_blocks(freq, in_len, ctx=gpu_ctx)
    n_dyn = 0
    for i in range(freq.shape[0]):
        f = freq[i].copy(); f[256] = 1
        ll, lc = _oracle_tree(oracle, 0, f[:286])
        dl, dc = _oracle_tree(oracle, 1, f[288:318])
        assert typ[i] in (1, 2)
        if typ[i] != 2:
            continue    # the static tree won: fixed codes, checked by the round-trip tests
        n_dyn += 1
        got_ll, got_lc = code[i, :286] >> 16, code[i, :286] & 0xffff
        got_dl, got_dc = code[i, 288:318] >> 16, code[i, 288:318] & 0xffff
        assert np.array_equal(got_ll, ll), f"case {i}: literal/length code lengths differ"
        assert np.array_equal(got_dl, dl), f"case {i}: distance code lengths differ"
        used = ll != 0
        assert np.array_equal(got_lc[used], lc[used]), f"case {i}: literal/length codes differ"
        used = dl != 0
        assert np.array_equal(got_dc[used], dc[used]), f"case {i}: distance codes differ"
    assert n_dyn >= 30


def test_block_type_choice(gpu_ctx):
    """_tr_flush_block (trees.ts:544-583): stored when the raw bytes are not longer than the best
    tree, static when it ties with dynamic."""
    B = pkg("batch")
    f = np.zeros((3, 320), np.uint32)
    f[0, :256] = 64                      # 16384 uniform literals: 8 bits each, header makes dynamic lose
    f[1, 101] = 3                        # a tiny block: static wins
    f[2, :256] = np.arange(256) % 7 * 30 + 1
    in_len = np.array([16384, 3, 1 << 20], np.uint32)
    code, typ, bits = B.huffman_blocks(f, in_len, ctx=gpu_ctx)
    assert typ[0] == 0 and bits[0] == 3 + 32 + 8 * 16384
    assert typ[1] == 1 and bits[1] == 3 + 3 * 8 + 7
    assert typ[2] == 2
