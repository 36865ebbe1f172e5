"""CPU suite: the C-ABI library loads and exports every symbol include/zsgpu.h declares."""
import ctypes
import os
import re

from conftest import ROOT, pkg


def test_library_exports_header_symbols():
    capi = pkg("capi")
    assert os.path.exists(capi.LIB_PATH), "build libzsgpu.so first (python __graft_entry__.py build)"
    lib = ctypes.CDLL(capi.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "zsgpu.h")).read()
    declared = set(re.findall(r"^ZS_API [^;(]*?\b(zs_[a-z0-9_]+)\(", header, flags=re.M))
    assert len(declared) >= 29
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in zsgpu.h but not exported"
    assert declared == set(capi.EXPORTED_SYMBOLS)


def test_no_cpu_fallback_without_gpu():
    import torch
    capi = pkg("capi")
    lib = capi.load()
    assert lib.zs_version().startswith(b"zsgpu")
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        rc = lib.zs_ctx_create(0, None, ctypes.byref(h))
        assert rc == capi.ZS_E_CUDA and not h.value          # fails loudly, no fallback


def test_host_side_combine_and_bounds():
    import zlib
    capi = pkg("capi")
    lib = capi.load()
    a, b = b"hello, ", b"world" * 1000
    assert lib.zs_crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
    assert lib.zs_adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)
    # deflateBound for the default state, deflate.ts:672
    for n in (0, 1, 1000, 1 << 20):
        assert lib.zs_deflate_bound(n, 0) == n + (n >> 12) + (n >> 14) + (n >> 25) + 7
        assert lib.zs_deflate_bound(n, 1) == lib.zs_deflate_bound(n, 0) + 6
        assert lib.zs_deflate_bound(n, 2) == lib.zs_deflate_bound(n, 0) + 18
    assert lib.zs_inflate_message(17) == b"invalid distance too far back"


def test_pointer_of_an_empty_view_keeps_its_position():
    """torch reports data_ptr() == 0 for an empty view; the engine needs where the view sits (the history
    of an empty part of a sharded stream lies before it)."""
    import torch
    B = pkg("batch")
    t = torch.arange(100, dtype=torch.uint8)
    v = t[60:60]
    assert v.data_ptr() == 0
    assert B._ptr(v).value == t.data_ptr() + 60
    assert B._ptr(t[10:20]).value == t.data_ptr() + 10
    assert B._ptr(None) is None
