// zs_common.cuh -- device helpers shared by the sm_100a kernels of libzsgpu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/zsgpu.h"

#define ZS_FULL_MASK 0xffffffffu

// ---- context (host side) -------------------------------------------------------------------------
struct zs_scratch {
    void* p = nullptr;
    size_t cap = 0;
};

struct zs_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    char err[256] = {0};
    uint64_t launches = 0;
    int sm_count = 148;
    // grow-only device scratch, one slot per purpose (see zs_api.cu)
    zs_scratch scr[40];
    // pinned host staging for small results
    void* h_pin = nullptr;
    size_t h_pin_cap = 0;
    // last inflate details (device ptr into scratch) for zs_inflate_last_details
    int32_t* d_last_detail = nullptr;
    uint32_t last_detail_n = 0;
    // copy streams + events of the pipelined host-buffer path (zs_deflate_batch)
    cudaStream_t s_in = nullptr, s_out = nullptr, s_res = nullptr;   // host-buffer path: H2D, bulk D2H, small result readbacks
    // streaming shim -> zs_inflate_batch (one stream): resume request and the block mark that comes back
    bool inflate_resume = false;
    uint64_t inflate_start_bit = 0;
    uint64_t inflate_mark[2] = {0, 0};
    uint64_t h_out_gen = 0;                      // bumped whenever the host-path output scratch is handed out again
    const uint8_t* last_inflate_out = nullptr;   // device copy of the output of the last one-stream zs_inflate_batch (valid until the next call)
    uint32_t seg_hint = 0;   // segment size (chunks) forced for the next deflate calls (slices of one batch)
    uint32_t inflate_batch_n = 0;   // streams of the whole batch while its slices are decoded (kernel choice), 0 = not sliced
    // one-stream inflate with sizes known on the host: lets zs_inflate_batch_dev use the segment-parallel decoder
    struct { bool on = false; uint64_t in_off = 0, in_len = 0, out_off = 0, out_cap = 0, dict_off = 0, dict_len = 0; } par;
    cudaEvent_t ev[64] = {nullptr};
    // profiling (zs_ctx_profile)
    bool prof_on = false;
    void* prof = nullptr;  // zs_profile*, zs_api.cu
};

void* zs_scratch_get(zs_ctx* ctx, int slot, size_t bytes);  // nullptr on failure (err set)
int zs_set_cuda_error(zs_ctx* ctx, cudaError_t e, const char* where);
#define ZS_CUDA_TRY(ctx, expr)                                                   \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess) return zs_set_cuda_error((ctx), _e, #expr);       \
    } while (0)
#ifdef ZS_DEBUG_HOOKS   // wait for every kernel so that a fault names the kernel that caused it
#define ZS_DEBUG_SYNC(ctx, e) do { if ((e) == cudaSuccess) (e) = cudaStreamSynchronize((ctx)->stream); } while (0)
#else
#define ZS_DEBUG_SYNC(ctx, e) do { } while (0)
#endif
#define ZS_LAUNCH_CHECK(ctx, name)                                               \
    do {                                                                         \
        (ctx)->launches++;                                                       \
        cudaError_t _e = cudaGetLastError();                                     \
        ZS_DEBUG_SYNC(ctx, _e);                                                  \
        if (_e != cudaSuccess) return zs_set_cuda_error((ctx), _e, name);        \
    } while (0)

// per-kernel CUDA-event timing (zs_ctx_profile): begin/end bracket one launch on ctx->stream
void zs_prof_begin(zs_ctx* ctx, const char* name);
void zs_prof_end(zs_ctx* ctx);
#define ZS_KERNEL(ctx, name, ...)        \
    do {                                 \
        zs_prof_begin((ctx), name);      \
        __VA_ARGS__;                     \
        zs_prof_end((ctx));              \
        ZS_LAUNCH_CHECK((ctx), name);    \
    } while (0)

// ---- device helpers --------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ unsigned zs_lane() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned zs_lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Unaligned little-endian loads from a read-only buffer whose base is 8-byte aligned.  `safe_end`
// is the buffer length rounded up to 8: aligned words below it are readable (cudaMalloc and the
// torch caching allocator hand out >= 256-byte granules), words at or above it read as zero.
__device__ __forceinline__ uint64_t zs_ld64(const uint8_t* __restrict__ base, uint64_t pos, uint64_t safe_end) {
    uint64_t a = pos & ~7ull;
    unsigned sh = (unsigned)(pos & 7u) * 8u;
    uint64_t lo = (a < safe_end) ? __ldg(reinterpret_cast<const unsigned long long*>(base + a)) : 0ull;
    if (sh == 0) return lo;
    uint64_t hi = (a + 8 < safe_end) ? __ldg(reinterpret_cast<const unsigned long long*>(base + a + 8)) : 0ull;
    return (lo >> sh) | (hi << (64u - sh));
}

__device__ __forceinline__ uint32_t zs_ld32(const uint8_t* __restrict__ base, uint64_t pos, uint64_t safe_end) {
    uint64_t a = pos & ~3ull;
    unsigned sh = (unsigned)(pos & 3u) * 8u;
    uint32_t lo = (a < safe_end) ? __ldg(reinterpret_cast<const unsigned*>(base + a)) : 0u;
    uint32_t hi = (sh != 0 && a + 4 < safe_end) ? __ldg(reinterpret_cast<const unsigned*>(base + a + 4)) : 0u;
    return __funnelshift_r(lo, hi, sh);
}

__device__ __forceinline__ unsigned zs_bitrev(unsigned code, unsigned len) { return __brev(code) >> (32u - len); }

// ---- RFC 1951 symbol geometry (deflate/constants.ts, trees-util.ts; regenerated from the RFC) ----
// length code index 0..28 for (match length - 3)
__device__ __forceinline__ unsigned zs_len_code(unsigned lc) {
    if (lc < 8) return lc;
    if (lc == 255) return 28;
    unsigned hb = 31u - __clz(lc);          // 3..7
    return ((hb - 1u) << 2) + ((lc >> (hb - 2u)) & 3u);
}
__device__ __forceinline__ unsigned zs_len_xbits(unsigned code) { return (code < 8 || code == 28) ? 0u : (code - 4u) >> 2; }
__device__ __forceinline__ unsigned zs_len_base(unsigned code) {  // base of (length - 3)
    if (code < 8) return code;
    if (code == 28) return 255;
    unsigned xb = (code - 4u) >> 2;
    return ((4u + (code & 3u)) << xb);
}
// distance code 0..29 for (distance - 1)
__device__ __forceinline__ unsigned zs_dist_code(unsigned d) {
    if (d < 4) return d;
    unsigned hb = 31u - __clz(d);           // >= 2
    return (hb << 1) + ((d >> (hb - 1u)) & 1u);
}
__device__ __forceinline__ unsigned zs_dist_xbits(unsigned code) { return code < 4 ? 0u : (code - 2u) >> 1; }
__device__ __forceinline__ unsigned zs_dist_base(unsigned code) {  // base of (distance - 1)
    if (code < 4) return code;
    unsigned xb = (code - 2u) >> 1;
    return (2u + (code & 1u)) << xb;
}

#endif  // __CUDACC__

// ---- kernel entry points (host wrappers, one per .cu) ---------------------------------------------
// checksum
int zs_launch_checksum_segments(zs_ctx* ctx, int kind, const uint8_t* d_buf, const uint64_t* d_off,
                                const uint64_t* d_len, uint32_t n, uint32_t* d_out);
int zs_launch_checksum_whole(zs_ctx* ctx, int kind, const uint8_t* d_buf, uint64_t len, uint32_t init,
                             uint32_t* d_result);
int zs_launch_checksum_stream(zs_ctx* ctx, int kind, const uint8_t* d_buf, const uint64_t* d_base, const uint64_t* d_len,
                              uint64_t max_len, uint32_t* d_result);
int zs_launch_checksum_fold(zs_ctx* ctx, int kind, const uint32_t* d_part, const uint64_t* d_off, uint32_t n,
                            uint32_t init, uint32_t* d_result);
uint32_t zs_host_crc32_combine(uint32_t c1, uint32_t c2, uint64_t len2);
uint32_t zs_host_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2);

// inflate
struct zs_inflate_args {
    const uint8_t* d_in;
    const uint64_t* d_in_off;
    uint32_t n;
    int wrap;       // bit0 zlib, bit1 gzip (both = auto)
    int deflate64;
    uint8_t* d_out;
    const uint64_t* d_out_off;
    uint64_t* d_out_len;
    uint64_t* d_in_used;
    int32_t* d_status;
    int32_t* d_detail;
    uint32_t* d_trailer;  // [2n]: stored check, stored isize (wrapped streams)
    uint32_t* d_flags;    // [n]: 1 = has zlib trailer (adler), 2 = has gzip trailer (crc)
    const uint8_t* d_dict;
    const uint64_t* d_dict_rng;
    int force_tps;        // tests: 1 = thread-per-stream kernel for any batch of >= 32 streams, -1 = never
    // streaming shim (warp kernel only): decode raw blocks from this bit of each stream instead of
    // from its start, and report where the last block that was begun starts: [2n] = (bit offset in the
    // stream, bytes produced before it) -- the point a later call can resume from
    const uint64_t* d_start_bit;
    uint64_t* d_block_mark;
    // segment-parallel decode of one stream (zs_inflate_par.cu):
    //  d_hdr_state != nullptr: parse the wrapper header only and report [5] = bit position of the first block
    //  (~0 if the header does not parse), [6] = trailer kind (1 zlib, 2 gzip)
    //  d_resume != nullptr: [4n] per stream = (start bit or ~0 for "from the beginning", bytes already
    //  produced, trailer kind | 4 = all blocks are decoded, -): continue behind what the parallel part did
    uint64_t* d_hdr_state;
    const uint64_t* d_resume;
};
int zs_launch_inflate(zs_ctx* ctx, const zs_inflate_args& a);
int zs_launch_inflate_parallel(zs_ctx* ctx, const uint8_t* d_in, uint64_t in_len, uint8_t* d_out, uint64_t out_cap,
                               const uint8_t* d_hist, uint64_t hist_len, uint64_t* d_state, uint64_t* d_resume);
int zs_launch_inflate_verify(zs_ctx* ctx, uint32_t n, const uint32_t* d_adler, const uint32_t* d_crc,
                             const uint32_t* d_trailer, const uint32_t* d_flags, const uint64_t* d_out_len,
                             uint32_t* d_checks, int32_t* d_status, int32_t* d_detail);

// deflate
struct zs_deflate_plan {
    const uint8_t* d_in;      // points at the first byte of this call's input
    uint64_t in_len;
    uint32_t history;         // readable bytes before d_in
    const uint64_t* d_in_off; // [n_chunks+1], device (always materialised)
    uint32_t n_chunks;
    uint32_t max_chunk;
    uint32_t max_bpc;         // block descriptor slots per chunk
    uint32_t seg_hint;        // 0 or the LZ77 segment size to use (chunks)
    int level, wrap, mode;
    int strategy;             // ZS_STRATEGY_*
    uint32_t flags;
    // scratch
    uint32_t* d_sym;          // [in_len] packed symbols
    uint32_t* d_chunk_nblk;   // [n_chunks]
    uint32_t* d_blk_desc;     // [n_chunks*max_bpc][4]: sym_begin(rel chunk), sym_count, in_begin(rel chunk), in_len
    uint32_t* d_blk_freq;     // [n_chunks*max_bpc][320]: 286 lit/len + 30 dist (+pad)
    uint32_t* d_blk_code;     // [n_chunks*max_bpc][320]: code | len<<16
    uint32_t* d_blk_hdr;      // [n_chunks*max_bpc][160]: word0 = type | hdr_bits<<8, words 1.. = header bits
    uint64_t* d_blk_bits;     // [n_chunks*max_bpc]: body size in bits (excl. stored padding) / abs bit offset
    uint64_t* d_chunk_pr;     // [n_chunks][2]: P (bits before first stored block, or all), R (rest), packed
    uint32_t* d_seg_counter;  // dynamic segment scheduler
    // outputs
    uint8_t* d_out;
    uint64_t out_cap;
    uint64_t* d_out_off;      // [n_chunks+1]
    uint64_t* d_out_bits;     // [n_chunks]
    uint32_t* d_checks;       // [n_chunks] or nullptr
    uint32_t* d_check_total;  // 1
    zs_deflate_result* d_result;
    int32_t* d_error;         // device flag: ZS_BUF_ERROR if out_cap too small
};
int zs_launch_lz77(zs_ctx* ctx, const zs_deflate_plan& p);
int zs_launch_huffman(zs_ctx* ctx, const zs_deflate_plan& p);
int zs_launch_bit_concat(zs_ctx* ctx, uint8_t* d_dst, uint64_t dst_bit_off, const uint8_t* d_src, uint64_t n_bits);
