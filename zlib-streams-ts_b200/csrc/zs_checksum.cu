// zs_checksum.cu -- K8: adler32 / crc32 over device buffers with combine-based parallel reduction.
//
// Results must be bit-exact with the reference's adler32 (src/mod/common/adler32.ts:4) and crc32
// (src/mod/common/crc32.ts:26).  The reference walks the buffer serially; here a warp owns one
// segment, every lane folds a contiguous slice (16-byte vector loads, slicing-by-16 tables in
// shared memory for crc32, dp4a byte sums for adler32) and the 32 lane results are merged with the
// combine algebra (crc: multiply by x^(8*bytes_after) mod P; adler: b += a * bytes_after).
// HBM-bound by design: algorithmic traffic = 1 read of every input byte.
#include <cstdio>

#include "zs_common.cuh"

namespace {

constexpr uint32_t kPoly = 0xedb88320u;
constexpr uint32_t kAdlerBase = 65521u;

// x^(8 * 2^k) mod P for k = 0..47 (reflected representation), filled by the host at start-up.
__constant__ uint32_t c_xpow[48];
// slicing-by-16 tables, built on the host once and kept in global memory.
__device__ uint32_t g_crc_tab[16][256];

// ---- GF(2)[x] mod P helpers (host + device) -----------------------------------------------------
__host__ __device__ inline uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (;;) {
        if (a & 0x80000000u) p ^= b;
        a <<= 1;
        if (a == 0) break;
        b = (b & 1u) ? ((b >> 1) ^ kPoly) : (b >> 1);
    }
    return p;
}

__device__ inline uint32_t dev_xpow8n(uint64_t nbytes) {
    uint32_t r = 0x80000000u;  // x^0
    for (int k = 0; nbytes != 0 && k < 48; ++k, nbytes >>= 1)
        if (nbytes & 1u) r = gf2_mulmod(r, c_xpow[k]);
    return r;
}

inline uint32_t host_xpow8n(uint64_t nbytes) {
    uint32_t r = 0x80000000u, sq = 0x00800000u;  // x^0, x^8
    while (nbytes) {
        if (nbytes & 1u) r = gf2_mulmod(r, sq);
        sq = gf2_mulmod(sq, sq);
        nbytes >>= 1;
    }
    return r;
}

__device__ inline uint32_t dev_crc_combine(uint32_t c1, uint32_t c2, uint64_t len2) {
    return len2 ? (gf2_mulmod(dev_xpow8n(len2), c1) ^ c2) : c1;
}
__host__ __device__ inline uint32_t adler_combine(uint32_t a1v, uint32_t a2v, uint64_t len2) {
    uint32_t rem = (uint32_t)(len2 % kAdlerBase);
    uint32_t a1 = a1v & 0xffffu, b1 = a1v >> 16, a2 = a2v & 0xffffu, b2 = a2v >> 16;
    uint32_t a = (a1 + a2 + kAdlerBase - 1u) % kAdlerBase;
    uint32_t b = (uint32_t)(((uint64_t)rem * ((a1 + kAdlerBase - 1u) % kAdlerBase) + b1 + b2) % kAdlerBase);
    return (b << 16) | a;
}

// ---- per-lane slice folds -------------------------------------------------------------------------
// raw crc remainder (state starts at `c`, no final xor) over [p, p+n); T = smem tables [16][256]
__device__ inline uint32_t crc_slice(const uint8_t* __restrict__ p, uint64_t n, uint32_t c,
                                     const uint32_t (*T)[256]) {
    // head: reach 16-byte alignment
    while (n && (reinterpret_cast<uintptr_t>(p) & 15u)) {
        c = T[0][(c ^ *p) & 0xffu] ^ (c >> 8);
        ++p; --n;
    }
    while (n >= 16) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        uint32_t a = v.x ^ c;
        c = T[15][a & 0xff] ^ T[14][(a >> 8) & 0xff] ^ T[13][(a >> 16) & 0xff] ^ T[12][a >> 24] ^
            T[11][v.y & 0xff] ^ T[10][(v.y >> 8) & 0xff] ^ T[9][(v.y >> 16) & 0xff] ^ T[8][v.y >> 24] ^
            T[7][v.z & 0xff] ^ T[6][(v.z >> 8) & 0xff] ^ T[5][(v.z >> 16) & 0xff] ^ T[4][v.z >> 24] ^
            T[3][v.w & 0xff] ^ T[2][(v.w >> 8) & 0xff] ^ T[1][(v.w >> 16) & 0xff] ^ T[0][v.w >> 24];
        p += 16; n -= 16;
    }
    while (n) {
        c = T[0][(c ^ *p) & 0xffu] ^ (c >> 8);
        ++p; --n;
    }
    return c;
}

// adler partial sums over [p, p+n): a = sum(bytes) mod BASE, b = sum((n - j) * byte_j) mod BASE
__device__ inline void adler_slice(const uint8_t* __restrict__ p, uint64_t n, uint32_t& a_out, uint32_t& b_out) {
    uint32_t a = 0, b = 0;
    while (n && (reinterpret_cast<uintptr_t>(p) & 15u)) {
        a += *p; b += a;
        ++p; --n;
    }
    a %= kAdlerBase; b %= kAdlerBase;
    while (n >= 16) {
        uint64_t blk = n < 2048 ? (n & ~15ull) : 2048;  // bytes before the next reduction
        for (uint64_t k = 0; k < blk; k += 16) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(p + k));
            // b += 16*a + 16*b0 + 15*b1 + ... + 1*b15 ; a += sum
            b += a << 4;
            b = __dp4a(v.x, 0x0d0e0f10u, b);
            b = __dp4a(v.y, 0x090a0b0cu, b);
            b = __dp4a(v.z, 0x05060708u, b);
            b = __dp4a(v.w, 0x01020304u, b);
            a = __dp4a(v.x, 0x01010101u, a);
            a = __dp4a(v.y, 0x01010101u, a);
            a = __dp4a(v.z, 0x01010101u, a);
            a = __dp4a(v.w, 0x01010101u, a);
        }
        a %= kAdlerBase; b %= kAdlerBase;
        p += blk; n -= blk;
    }
    while (n) {
        a += *p; b += a;
        ++p; --n;
    }
    a_out = a % kAdlerBase;
    b_out = b % kAdlerBase;
}

// One warp per segment, grid-stride.  KIND 0 adler32, 1 crc32.
template <int KIND>
__global__ void __launch_bounds__(256) checksum_segments_kernel(const uint8_t* __restrict__ buf,
                                                                const uint64_t* __restrict__ off,
                                                                const uint64_t* __restrict__ seg_len, uint32_t n,
                                                                uint32_t* __restrict__ out) {
    __shared__ uint32_t T[KIND ? 16 : 1][256];
    if (KIND) {
        for (unsigned i = threadIdx.x; i < 16 * 256; i += blockDim.x) T[i >> 8][i & 255] = g_crc_tab[i >> 8][i & 255];
        __syncthreads();
    }
    const unsigned lane = zs_lane();
    const unsigned warps_per_cta = blockDim.x >> 5;
    for (uint64_t seg = (uint64_t)blockIdx.x * warps_per_cta + (threadIdx.x >> 5); seg < n;
         seg += (uint64_t)gridDim.x * warps_per_cta) {
        // segment = [off[seg], off[seg+1]) or, when seg_len is given, seg_len[seg] bytes at off[seg]
        const uint64_t beg = off[seg];
        const uint64_t len = seg_len ? seg_len[seg] : off[seg + 1] - beg;
        // slice length: multiple of 16 so that interior slices stay vector aligned
        uint64_t S = ((len + 31) / 32 + 15) & ~15ull;
        if (S == 0) S = 16;
        uint64_t sb = (uint64_t)lane * S, se = sb + S;
        if (sb > len) sb = len;
        if (se > len) se = len;
        const uint64_t after = len - se;
        if (KIND) {
            uint32_t r = crc_slice(buf + beg + sb, se - sb, 0u, T);
            // the 0xffffffff preset behaves like a prefix term shifted over the whole segment
            if (lane == 0) r ^= gf2_mulmod(dev_xpow8n(se - sb), 0xffffffffu);
            if (after) r = gf2_mulmod(dev_xpow8n(after), r);
            for (int o = 16; o; o >>= 1) r ^= __shfl_xor_sync(ZS_FULL_MASK, r, o);
            if (lane == 0) out[seg] = ~r;
        } else {
            uint32_t a, b;
            adler_slice(buf + beg + sb, se - sb, a, b);
            uint64_t bb = (uint64_t)b + (uint64_t)a * (after % kAdlerBase);
            uint32_t b2 = (uint32_t)(bb % kAdlerBase);
            for (int o = 16; o; o >>= 1) {
                a += __shfl_xor_sync(ZS_FULL_MASK, a, o);
                b2 += __shfl_xor_sync(ZS_FULL_MASK, b2, o);
            }
            if (lane == 0) {
                uint32_t A = (1u + a) % kAdlerBase;
                uint32_t B = (uint32_t)(((uint64_t)b2 + len % kAdlerBase) % kAdlerBase);
                out[seg] = (B << 16) | A;
            }
        }
    }
}

// Fold n per-segment checksums (segment i covers off[i+1]-off[i] bytes) into one value continuing
// from `init`.  Single CTA; each thread folds a contiguous run, then a shared-memory tree.
template <int KIND>
__global__ void __launch_bounds__(1024) checksum_fold_kernel(const uint32_t* __restrict__ part,
                                                             const uint64_t* __restrict__ off, uint32_t n,
                                                             uint32_t init, uint32_t* __restrict__ result) {
    __shared__ uint32_t s_val[1024];
    __shared__ uint64_t s_len[1024];
    const unsigned t = threadIdx.x, nt = blockDim.x;
    const uint32_t per = (n + nt - 1) / nt;
    uint32_t b = t * per, e = b + per;
    if (b > n) b = n;
    if (e > n) e = n;
    uint32_t acc = KIND ? 0u : 1u;
    uint64_t alen = 0;
    for (uint32_t i = b; i < e; ++i) {
        uint64_t l = off[i + 1] - off[i];
        acc = KIND ? dev_crc_combine(acc, part[i], l) : adler_combine(acc, part[i], l);
        alen += l;
    }
    s_val[t] = acc;
    s_len[t] = alen;
    __syncthreads();
    for (unsigned stride = 1; stride < nt; stride <<= 1) {
        if ((t & (2 * stride - 1)) == 0 && t + stride < nt) {
            uint32_t v2 = s_val[t + stride];
            uint64_t l2 = s_len[t + stride];
            s_val[t] = KIND ? dev_crc_combine(s_val[t], v2, l2) : adler_combine(s_val[t], v2, l2);
            s_len[t] += l2;
        }
        __syncthreads();
    }
    if (t == 0) {
        uint32_t v = s_val[0];
        uint64_t l = s_len[0];
        *result = KIND ? dev_crc_combine(init, v, l) : adler_combine(init, v, l);
    }
}

__global__ void make_offsets_kernel(uint64_t* off, uint64_t len, uint64_t piece, uint32_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) {
        uint64_t v = i * piece;
        off[i] = v < len ? v : len;
    }
}

bool g_tables_ready[64] = {false};

int ensure_tables(zs_ctx* ctx) {
    if (ctx->device < 64 && g_tables_ready[ctx->device]) return ZS_OK;
    static uint32_t tab[16][256];
    for (unsigned n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (kPoly ^ (c >> 1)) : (c >> 1);
        tab[0][n] = c;
    }
    for (unsigned n = 0; n < 256; n++)
        for (int k = 1; k < 16; k++) {
            uint32_t p = tab[k - 1][n];
            tab[k][n] = (p >> 8) ^ tab[0][p & 0xffu];
        }
    uint32_t xp[48];
    uint32_t sq = 0x00800000u;
    for (int k = 0; k < 48; k++) {
        xp[k] = sq;
        sq = gf2_mulmod(sq, sq);
    }
    ZS_CUDA_TRY(ctx, cudaMemcpyToSymbolAsync(g_crc_tab, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA_TRY(ctx, cudaMemcpyToSymbolAsync(c_xpow, xp, sizeof(xp), 0, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // `tab`/`xp` are host temporaries
    if (ctx->device < 64) g_tables_ready[ctx->device] = true;
    return ZS_OK;
}

}  // namespace

uint32_t zs_host_crc32_combine(uint32_t c1, uint32_t c2, uint64_t len2) {
    return len2 ? (gf2_mulmod(host_xpow8n(len2), c1) ^ c2) : c1;
}
uint32_t zs_host_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2) { return adler_combine(a1, a2, len2); }

int zs_launch_checksum_segments(zs_ctx* ctx, int kind, const uint8_t* d_buf, const uint64_t* d_off,
                                const uint64_t* d_len, uint32_t n, uint32_t* d_out) {
    if (n == 0) return ZS_OK;
    int rc = ensure_tables(ctx);
    if (rc != ZS_OK) return rc;
    unsigned ctas = (n + 7) / 8;
    unsigned cap = (unsigned)ctx->sm_count * 8u;
    if (ctas > cap) ctas = cap;
    if (kind)
        ZS_KERNEL(ctx, "checksum_segments_kernel", checksum_segments_kernel<1><<<ctas, 256, 0, ctx->stream>>>(d_buf, d_off, d_len, n, d_out));
    else
        ZS_KERNEL(ctx, "checksum_segments_kernel", checksum_segments_kernel<0><<<ctas, 256, 0, ctx->stream>>>(d_buf, d_off, d_len, n, d_out));
    return ZS_OK;
}

int zs_launch_checksum_fold(zs_ctx* ctx, int kind, const uint32_t* d_part, const uint64_t* d_off, uint32_t n,
                            uint32_t init, uint32_t* d_result) {
    int rc = ensure_tables(ctx);
    if (rc != ZS_OK) return rc;
    if (kind)
        ZS_KERNEL(ctx, "checksum_fold_kernel", checksum_fold_kernel<1><<<1, 1024, 0, ctx->stream>>>(d_part, d_off, n, init, d_result));
    else
        ZS_KERNEL(ctx, "checksum_fold_kernel", checksum_fold_kernel<0><<<1, 1024, 0, ctx->stream>>>(d_part, d_off, n, init, d_result));
    return ZS_OK;
}

// Whole-buffer checksum: cut into 64 KiB pieces (one warp each), then fold.
int zs_launch_checksum_whole(zs_ctx* ctx, int kind, const uint8_t* d_buf, uint64_t len, uint32_t init,
                             uint32_t* d_result) {
    const uint64_t piece = 65536;
    uint64_t n64 = (len + piece - 1) / piece;
    if (n64 == 0) n64 = 1;
    if (n64 > 0xfffffff0ull) {
        snprintf(ctx->err, sizeof(ctx->err), "checksum: buffer too large");
        return ZS_STREAM_ERROR;
    }
    uint32_t n = (uint32_t)n64;
    uint64_t* d_off = (uint64_t*)zs_scratch_get(ctx, 0, (size_t)(n + 1) * sizeof(uint64_t));
    uint32_t* d_part = (uint32_t*)zs_scratch_get(ctx, 1, (size_t)n * sizeof(uint32_t));
    if (!d_off || !d_part) return ZS_MEM_ERROR;
    ZS_KERNEL(ctx, "make_offsets_kernel", make_offsets_kernel<<<(n + 1 + 255) / 256, 256, 0, ctx->stream>>>(d_off, len, piece, n));
    int rc = zs_launch_checksum_segments(ctx, kind, d_buf, d_off, nullptr, n, d_part);
    if (rc != ZS_OK) return rc;
    return zs_launch_checksum_fold(ctx, kind, d_part, d_off, n, init, d_result);
}
