// zs_stream.cu -- placeholder, replaced by the streaming shim
#include "zs_common.cuh"
extern "C" {
int zs_stream_deflate_init(zs_ctx*, zs_stream*, int, int, int, int, int) { return ZS_STREAM_ERROR; }
int zs_stream_deflate_set_dictionary(zs_stream*, const uint8_t*, uint32_t) { return ZS_STREAM_ERROR; }
int zs_stream_deflate(zs_stream*, int) { return ZS_STREAM_ERROR; }
int zs_stream_deflate_end(zs_stream*) { return ZS_STREAM_ERROR; }
int zs_stream_inflate_init(zs_ctx*, zs_stream*, int) { return ZS_STREAM_ERROR; }
int zs_stream_inflate_set_dictionary(zs_stream*, const uint8_t*, uint32_t) { return ZS_STREAM_ERROR; }
int zs_stream_inflate(zs_stream*, int) { return ZS_STREAM_ERROR; }
int zs_stream_inflate_reset(zs_stream*) { return ZS_STREAM_ERROR; }
int zs_stream_inflate_end(zs_stream*) { return ZS_STREAM_ERROR; }
}
