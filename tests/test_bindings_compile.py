"""The Node-API addon of INTEGRATION.md cannot run here (no node), but it must at least type-check
against include/zsgpu.h and bind only symbols libzsgpu.so really exports."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADDON = os.path.join(ROOT, "bindings", "node", "zsgpu_addon.cc")
LIB = os.path.join(ROOT, "zlib-streams-ts_b200", "libzsgpu.so")


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_addon_compiles_and_binds_exported_symbols(tmp_path):
    obj = tmp_path / "addon.o"
    r = subprocess.run(
        ["g++", "-std=c++17", "-Wall", "-Werror", "-c", "-fPIC", "-I" + os.path.join(ROOT, "tests", "stubs"),
         "-I" + os.path.join(ROOT, "include"), ADDON, "-o", str(obj)],
        capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    need = set(re.findall(r"\bU (zs_\w+)", subprocess.run(["nm", "-u", str(obj)], capture_output=True, text=True).stdout))
    assert "zs_stream_inflate_get_header" in need and "zs_deflate_batch" in need
    if not os.path.exists(LIB):
        pytest.skip("libzsgpu.so not built")
    have = set(re.findall(r"\bT (zs_\w+)", subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True).stdout))
    assert need <= have, sorted(need - have)


def test_facade_exports_the_reference_names():
    """Every low-level function the reference's index files export and INTEGRATION.md lists as bound has
    an export of the same name in the TypeScript facade."""
    src = open(os.path.join(ROOT, "bindings", "ts", "zlib-streams-gpu.ts")).read()
    exported = set(re.findall(r"export (?:function|const) (\w+)", src))
    for m in re.finditer(r"export const ([^;]+);", src):
        exported |= set(re.findall(r"(\w+) =", m.group(1)))
    for name in ("createDeflateStream", "deflateInit", "deflateInit2_", "deflate", "deflateEnd", "deflateSetDictionary",
                 "deflateReset", "deflateResetKeep", "deflateParams", "deflatePending", "deflateUsed", "deflateSetHeader",
                 "createInflateStream", "inflateInit", "inflateInit2_", "inflate", "inflateEnd", "inflateReset",
                 "inflateReset2", "inflateSetDictionary", "inflateGetHeader", "adler32", "crc32",
                 "deflateBatch", "inflateBatch"):
        assert name in exported, name
