/* stream_pump.c -- the driver loop of the reference's CompressionStream / DecompressionStream in C.
 *
 * src/mod/streams.ts:78-93 (transform) and :139-170 (flush) drive deflate()/inflate() like this: every input
 * piece is handed over in slices of at most 32 KiB (streams.ts:7), each slice is passed with Z_NO_FLUSH -- the
 * last one of the stream with Z_FINISH -- and the call is repeated with a fresh 64 KiB output buffer for as long
 * as it fills the buffer or still holds input; whatever a call produced is copied out and enqueued.  A Node-API
 * addon costs about a microsecond per call, a Python/ctypes loop 30-50, so the throughput of the drop-in stream
 * API is measured with this loop (bench.py -> extra.stream_api, tests/test_zlib_api_gpu.py), not with a Python one.
 *
 * No dependency on libzsgpu at link time: the caller passes zs_stream_deflate or zs_stream_inflate.
 * Build: gcc -O2 -shared -fPIC -I include bindings/c/stream_pump.c -o bindings/c/_build/libstreampump.so */
#include <stdint.h>
#include <string.h>

#include "zsgpu.h"

typedef int (*zs_step_fn)(zs_stream*, int);

#ifdef __cplusplus
extern "C"
#endif
__attribute__((visibility("default")))
/* Runs the whole stream src[0, n) through `step`.  Output is appended to dst (capacity dst_cap).  Returns the
 * last return code of `step` (Z_STREAM_END = 1 on success); *produced = bytes written to dst, *calls = calls made.
 * -100 - rc: dst too small. */
int zs_stream_pump(zs_step_fn step, zs_stream* strm, const uint8_t* src, uint64_t n, uint64_t in_slice, uint8_t* dst,
                   uint64_t dst_cap, uint8_t* obuf, uint64_t out_slice, uint64_t* produced, uint64_t* calls) {
    uint64_t pos = 0, made_total = 0, ncalls = 0;
    int rc = ZS_OK;
    for (;;) {
        const uint64_t take = n - pos < in_slice ? n - pos : in_slice;
        strm->next_in = src + pos;
        strm->avail_in = take;
        pos += take;
        const int flush = pos >= n ? ZS_FINISH : ZS_NO_FLUSH;
        for (;;) {
            strm->next_out = obuf;
            strm->avail_out = out_slice;
            rc = step(strm, flush);
            ++ncalls;
            const uint64_t made = out_slice - strm->avail_out;
            if (made) {
                if (made_total + made > dst_cap) { *produced = made_total; *calls = ncalls; return -100 - rc; }
                memcpy(dst + made_total, obuf, made);
                made_total += made;
            }
            if (rc == ZS_STREAM_END) goto done;
            if (rc == ZS_BUF_ERROR && flush != ZS_FINISH) break;   /* nothing more to do with this slice */
            if (rc != ZS_OK) goto done;
            if (flush != ZS_FINISH && strm->avail_in == 0 && strm->avail_out != 0) break;
        }
    }
done:
    *produced = made_total;
    *calls = ncalls;
    return rc;
}
