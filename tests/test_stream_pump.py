"""CPU: the C driver loop of the stream API (bindings/c/stream_pump.c, what bench.py and the GPU tests drive
zs_stream_deflate / zs_stream_inflate with) against a mock codec that follows the z_stream protocol of
src/mod/deflate/deflate.ts:716-748 the way the GPU shim does: it buffers input, hands output out only in bursts,
answers Z_BUF_ERROR when a call can make no progress and Z_STREAM_END once everything has been delivered after
Z_FINISH.  Checks the loop of src/mod/streams.ts:78-93,139-170 as restated in C: slicing, draining through a small
output buffer, termination, the dst-too-small verdict."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, pkg

MOCK = r"""
#include <stdint.h>
#include <string.h>
#include "zsgpu.h"
/* identity "codec": input is buffered (at most 300000 bytes are held back), output appears in bursts */
static uint8_t held[1 << 22]; static uint64_t n_held, n_out, finished;
void mock_reset(void) { n_held = n_out = finished = 0; }
int mock_step(zs_stream* s, int flush) {
    if (!s->next_out || (s->avail_in && !s->next_in)) return ZS_STREAM_ERROR;
    if (s->avail_out == 0) return ZS_BUF_ERROR;
    const uint64_t in0 = s->avail_in, out0 = s->avail_out;
    memcpy(held + n_held, s->next_in, s->avail_in);
    n_held += s->avail_in; s->next_in += s->avail_in; s->total_in += s->avail_in; s->avail_in = 0;
    if (flush == ZS_FINISH) finished = 1;
    /* deliverable: everything once finished, otherwise whole bursts of 300000 bytes */
    uint64_t upto = finished ? n_held : (n_held / 300000) * 300000;
    uint64_t c = upto > n_out ? upto - n_out : 0;
    if (c > s->avail_out) c = s->avail_out;
    memcpy(s->next_out, held + n_out, c);
    n_out += c; s->next_out += c; s->avail_out -= c; s->total_out += c;
    if (finished && n_out == n_held) return ZS_STREAM_END;
    if (in0 == s->avail_in && out0 == s->avail_out) return ZS_BUF_ERROR;
    return ZS_OK;
}
"""


@pytest.fixture(scope="module")
def mock(tmp_path_factory):
    d = tmp_path_factory.mktemp("pump")
    src = d / "mock.c"
    src.write_text(MOCK)
    so = d / "libmock.so"
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(so)], check=True)
    return C.CDLL(str(so))


def test_pump_against_mock_codec(mock):
    from tools import streampump
    capi = pkg("capi")
    rng = np.random.default_rng(7)
    for n in (0, 1, 32768, 32769, 1000000, 3 << 20):
        data = rng.integers(0, 256, n, dtype=np.uint8)
        for in_slice, out_slice in ((32768, 65536), (1000, 7), (1 << 20, 4096)):
            if n > 1000000 and out_slice == 7:
                continue
            mock.mock_reset()
            zs = capi.ZStream()
            dst = np.zeros(n + 16, dtype=np.uint8)
            rc, made, calls = streampump.pump(mock.mock_step, zs, data, dst, in_slice, out_slice)
            assert rc == capi.Z_STREAM_END, (n, in_slice, out_slice, rc)
            assert made == n and zs.total_in == n and zs.total_out == n
            assert np.array_equal(dst[:n], data)
            assert calls >= max(1, -(-n // in_slice))
    # a destination that is too small is reported, not overrun
    mock.mock_reset()
    zs = capi.ZStream()
    data = rng.integers(0, 256, 500000, dtype=np.uint8)
    dst = np.zeros(100000, dtype=np.uint8)
    rc, made, calls = streampump.pump(mock.mock_step, zs, data, dst)
    assert rc <= -100 and made <= dst.size
