// zs_lz77.cu -- K1/K2/K3: LZ77 match finding (greedy levels 1-3, lazy levels 4-9), parse and symbol
// histogramming for independent or dictionary-primed chunks.
//
// Replaces, for whole chunks, the reference's per-byte loop: INSERT_STRING / hash chains
// (src/mod/deflate/deflate.ts:109-141), longest_match (:1053-1115), deflate_fast (:1281-1350),
// deflate_slow incl. the TOO_FAR rule (:1352-1448), _tr_tally_lit/_tr_tally_dist and d_code
// (src/mod/deflate/utils.ts:55-89), with CONFIGURATION_TABLE (deflate.ts:86-103) giving
// max_chain / nice_length / max_lazy per level.  The bit stream differs from the reference's (the
// north star allows it); what must hold is: the symbols decode to the input, and the compressed
// size stays within 3 % of the reference at the same level (tests/test_deflate_gpu.py).
//
// Work decomposition (B200: 148 SMs, 227 KB shared memory per CTA):
//   * one persistent CTA per SM pulls *segments* (runs of consecutive chunks) from an atomic
//     counter; the 2 x 64 KiB hash-chain tables (head[32768], prev[32768], 16-bit positions like the
//     reference's) live in shared memory and are carried from chunk to chunk inside a segment, so
//     dictionary priming costs one 32 KiB insert-only pass per segment, not per chunk;
//   * inside the CTA the per-byte serial loop of the reference is cut into three roles that run
//     concurrently on different 448-position steps, one __syncthreads per step:
//        warp 0  (insert) : hashes 32 positions per trip, links equal hashes inside the warp with
//                           __match_any_sync, threads the chains through head/prev in position order;
//        warps 2..15 (search): one lane per position walks that position's chain (<= max_chain
//                           candidates, 8-byte compares through L1) -- every position is searched
//                           speculatively, which is what makes the chain walk 448-wide;
//        warp 1  (parse)  : turns the per-position longest matches into the greedy / lazy parse with a
//                           5-round pointer-doubling over 32 positions, writes packed symbols,
//                           updates the block histogram (shared-memory atomics) and cuts blocks
//                           every <= 16383 symbols (lit_bufsize - 1, deflate.ts:323,336).
//     The inserter never runs more than 2 steps ahead of a searcher, so restricting matches to
//     32768 - 2*448 bytes keeps every prev[] slot a searcher reads stable (no races, deterministic).
//   * the window itself is read from HBM through L1/L2 (64-bit loads, unaligned handled by funnel
//     shifts); with 128 KiB of tables resident the remaining ~100 KB of L1 holds the 32 KiB window.
#include <cstdio>

#include "zs_common.cuh"

namespace {

constexpr int kSearchWarps = 14;
constexpr int kWarpInsert = 0;
constexpr int kWarpParse = 1;
constexpr int kThreads = 32 * (kSearchWarps + 2);
constexpr int kStep = 32 * kSearchWarps;               // positions per pipeline step
constexpr unsigned kMaxDist = 32768u - 2u * kStep;      // see header comment
constexpr unsigned kSymLimit = 16383u;                  // symbols per block, deflate.ts:336
constexpr unsigned kTooFar = 4096u;                     // deflate/constants.ts (TOO_FAR)
constexpr unsigned kHashBits = 15;

struct LevelCfg { int lazy_fn, good, lazy, nice, chain; };
// CONFIGURATION_TABLE, deflate.ts:86-103
__constant__ LevelCfg c_levels[10] = {
    {0, 0, 0, 0, 0},      {0, 4, 4, 8, 4},      {0, 4, 5, 16, 8},      {0, 4, 6, 32, 32},
    {1, 4, 4, 16, 16},    {1, 8, 16, 32, 32},   {1, 8, 16, 128, 128},  {1, 8, 32, 128, 256},
    {1, 32, 128, 258, 1024}, {1, 32, 258, 258, 4096}};

struct Smem {
    uint16_t head[1 << kHashBits];
    uint16_t prev[32768];
    uint32_t res[3 * kStep];  // per-position best match: len << 16 | dist (0 = none), 3-step ring
    uint32_t hist[320];       // 0..285 literal/length, 288..317 distance
    uint32_t seg;             // segment being processed
};

struct LzArgs {
    const uint8_t* buf;        // 16-byte aligned; position 0 of all indices below
    uint64_t org;              // index of the call's first input byte (d_in[0]) inside buf
    uint64_t valid_lo;         // first readable index (org - history)
    uint64_t data_end;         // one past the last input byte
    uint64_t safe_end;         // data_end rounded up to 8
    const uint64_t* in_off;    // [n_chunks+1] relative to org
    uint32_t n_chunks, seg_chunks, n_seg, max_bpc;
    int level;
    int cross;                 // 1: matches may reach before the chunk start (PRIME / STITCHED)
    uint32_t* sym;
    uint32_t* chunk_nblk;
    uint32_t* blk_desc;
    uint32_t* blk_freq;
    uint32_t* seg_counter;
};

__device__ __forceinline__ unsigned hash3(uint32_t w) { return ((w & 0xffffffu) * 0x9E3779B1u) >> (32 - kHashBits); }

// ---- insert: 32 consecutive positions, chains kept in position order ---------------------------
__device__ __forceinline__ void insert_batch(Smem& S, const LzArgs& a, uint64_t b, uint64_t limit, uint64_t lo) {
    const unsigned lane = zs_lane();
    const uint64_t p = b + lane;
    const bool valid = p < limit && p + 2 < a.data_end;
    unsigned h = 0x10000u + lane;
    if (valid) h = hash3(zs_ld32(a.buf, p, a.safe_end));
    const unsigned peers = __match_any_sync(ZS_FULL_MASK, h);
    if (valid) {
        const unsigned lower = peers & zs_lanemask_lt();
        uint64_t pred = p;  // "no predecessor" is encoded as a self link (delta 0)
        if (lower) {
            pred = b + (31u - __clz(lower));
        } else {
            unsigned delta = ((unsigned)p - S.head[h]) & 0xffffu;
            if (delta != 0 && delta <= kMaxDist && p >= lo + delta) pred = p - delta;
        }
        S.prev[p & 32767u] = (uint16_t)pred;
        if ((peers >> lane) == 1u) S.head[h] = (uint16_t)p;  // highest lane of the group
    }
    __syncwarp();
}

// ---- search: one lane per position -------------------------------------------------------------
__device__ __forceinline__ uint32_t search_position(const Smem& S, const LzArgs& a, const LevelCfg& cfg, uint64_t p,
                                                    uint64_t chunk_end, uint64_t lo) {
    uint64_t room = chunk_end - p;
    const unsigned max_len = room < 258 ? (unsigned)room : 258u;
    if (max_len < 3) return 0;
    const unsigned nice = (unsigned)cfg.nice < max_len ? (unsigned)cfg.nice : max_len;
    const uint64_t p0 = zs_ld64(a.buf, p, a.safe_end);
    unsigned best_len = 2, best_dist = 0;
    uint64_t cur = p;
    for (int chain = cfg.chain; chain > 0; --chain) {
        unsigned delta = ((unsigned)cur - S.prev[cur & 32767u]) & 0xffffu;
        if (delta == 0) break;
        if (cur < lo + delta) break;
        uint64_t cand = cur - delta;
        if (p - cand > kMaxDist) break;
        cur = cand;
        uint64_t x = zs_ld64(a.buf, cand, a.safe_end) ^ p0;
        if ((x & 0xffffffull) != 0) continue;  // hash collision
        if (best_len >= 8) {                    // cannot beat the best unless the bytes up to best_len agree
            uint64_t y = zs_ld64(a.buf, cand + best_len - 7, a.safe_end) ^ zs_ld64(a.buf, p + best_len - 7, a.safe_end);
            if (y) continue;
        }
        unsigned len;
        if (x) {
            len = (unsigned)(__ffsll((long long)x) - 1) >> 3;
        } else {
            len = 8;
            while (len < max_len) {
                uint64_t y = zs_ld64(a.buf, cand + len, a.safe_end) ^ zs_ld64(a.buf, p + len, a.safe_end);
                if (y) { len += (unsigned)(__ffsll((long long)y) - 1) >> 3; break; }
                len += 8;
            }
        }
        if (len > max_len) len = max_len;
        if (len > best_len) {
            best_len = len;
            best_dist = (unsigned)(p - cand);
            if (len >= nice) break;
        }
    }
    if (best_len < 3) return 0;
    if (cfg.lazy_fn && best_len == 3 && best_dist > kTooFar) return 0;  // deflate.ts:1381-1387
    return (best_len << 16) | best_dist;
}

// ---- parse state (kept by the parse warp, identical in all its lanes) ---------------------------
struct ParseState {
    uint32_t ppos;       // next batch start, relative to the chunk
    uint32_t skip;       // leading positions of that batch already covered by an emitted match
    uint32_t nsym;       // symbols emitted in the chunk
    uint32_t blk;        // blocks closed in the chunk
    uint32_t blk_sym0;   // first symbol of the open block
    uint32_t blk_pos0;   // first input position of the open block
};

__device__ __forceinline__ void close_block(Smem& S, const LzArgs& a, ParseState& ps, uint32_t chunk, uint32_t pos_end) {
    const unsigned lane = zs_lane();
    __syncwarp();
    const bool fits = ps.blk < a.max_bpc;  // max_bpc is sized so that this always holds
    const uint64_t bi = (uint64_t)chunk * a.max_bpc + (fits ? ps.blk : 0u);
    if (fits && lane == 0) {
        uint32_t* d = a.blk_desc + bi * 4;
        d[0] = ps.blk_sym0;
        d[1] = ps.nsym - ps.blk_sym0;
        d[2] = ps.blk_pos0;
        d[3] = pos_end - ps.blk_pos0;
    }
    uint32_t* f = a.blk_freq + bi * 320;
    for (unsigned i = lane; i < 320; i += 32) {
        if (fits) f[i] = S.hist[i];
        S.hist[i] = 0;
    }
    __syncwarp();
    ps.blk++;
    ps.blk_sym0 = ps.nsym;
    ps.blk_pos0 = pos_end;
}

// One batch of 32 positions [q0, q0+32) of chunk `chunk` (n bytes at absolute index cbase).
__device__ __forceinline__ void parse_batch(Smem& S, const LzArgs& a, const LevelCfg& cfg, ParseState& ps,
                                            uint32_t chunk, uint64_t cbase, uint32_t n, uint32_t q0) {
    const unsigned lane = zs_lane();
    if (ps.skip >= 32) { ps.skip -= 32; return; }
    const uint32_t q = q0 + lane;
    const bool in_range = q < n;
    uint32_t r = 0;
    if (in_range) r = S.res[((q / kStep) % 3) * kStep + (q % kStep)];
    unsigned L = r >> 16, D = r & 0xffffu;
    unsigned Ln = __shfl_down_sync(ZS_FULL_MASK, L, 1);
    if (lane == 31) {
        const uint32_t qn = q + 1;
        Ln = qn < n ? (S.res[((qn / kStep) % 3) * kStep + (qn % kStep)] >> 16) : 0u;
    }
    // deflate_slow's lazy evaluation (deflate.ts:1372-1426): the match at q is dropped for a
    // literal when the match at q+1 is strictly longer and L < max_lazy
    const bool deferred = cfg.lazy_fn && L >= 3 && L < (unsigned)cfg.lazy && Ln > L;
    const bool is_match = L >= 3 && !deferred;
    unsigned J = lane + (is_match ? L : 1u);
    unsigned visited, exitJ;
    if (__ballot_sync(ZS_FULL_MASK, is_match) == 0) {
        visited = 0xffffffffu << ps.skip;
        exitJ = 32;
    } else {
        unsigned M = 1u << lane;
#pragma unroll
        for (int round = 0; round < 5; ++round) {
            const unsigned src = J < 32 ? J : lane;
            const unsigned Mj = __shfl_sync(ZS_FULL_MASK, M, src);
            const unsigned Jj = __shfl_sync(ZS_FULL_MASK, J, src);
            if (J < 32) { M |= Mj; J = Jj; }
        }
        visited = __shfl_sync(ZS_FULL_MASK, M, ps.skip);
        exitJ = __shfl_sync(ZS_FULL_MASK, J, ps.skip);
    }
    if (n - q0 < 32) visited &= (1u << (n - q0)) - 1u;
    const unsigned nv = __popc(visited);
    if (ps.nsym - ps.blk_sym0 + nv > kSymLimit) {
        // the open block ends before this batch's first emitted position
        const unsigned first = __ffs(visited) - 1u;
        close_block(S, a, ps, chunk, q0 + first);
    }
    if ((visited >> lane) & 1u) {
        const unsigned idx = ps.nsym + __popc(visited & zs_lanemask_lt());
        uint32_t packed;
        if (is_match) {
            packed = (D << 16) | L;
            atomicAdd(&S.hist[257u + zs_len_code(L - 3u)], 1u);
            atomicAdd(&S.hist[288u + zs_dist_code(D - 1u)], 1u);
        } else {
            const unsigned byte = __ldg(a.buf + cbase + q);
            packed = byte;
            atomicAdd(&S.hist[byte], 1u);
        }
        a.sym[(cbase - a.org) + idx] = packed;
    }
    ps.nsym += nv;
    ps.skip = exitJ - 32u;
    if (n - q0 < 32) ps.skip = 0;
}

__global__ void __launch_bounds__(kThreads, 1) lz77_kernel(LzArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const unsigned wid = threadIdx.x >> 5, lane = zs_lane();
    const LevelCfg cfg = c_levels[a.level];

    for (;;) {
        if (threadIdx.x == 0) S.seg = atomicAdd(a.seg_counter, 1u);
        __syncthreads();
        const uint32_t seg = S.seg;
        if (seg >= a.n_seg) break;
        const uint32_t c0 = seg * a.seg_chunks;
        const uint32_t c1 = (c0 + a.seg_chunks < a.n_chunks) ? c0 + a.seg_chunks : a.n_chunks;

        // fresh tables per segment: the output never depends on which CTA ran which segment
        {
            uint4* z = reinterpret_cast<uint4*>(S.head);
            const unsigned nz = (sizeof(S.head) + sizeof(S.prev)) / sizeof(uint4);
            for (unsigned i = threadIdx.x; i < nz; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
            for (unsigned i = threadIdx.x; i < 320; i += kThreads) S.hist[i] = 0;
        }
        __syncthreads();

        // dictionary priming (deflateSetDictionary, deflate.ts:367-424): insert-only pass over the
        // <= 32 KiB that precede the segment
        const uint64_t seg_start = a.org + a.in_off[c0];
        if (a.cross && wid == kWarpInsert) {
            uint64_t ps0 = seg_start > a.valid_lo + 32768 ? seg_start - 32768 : a.valid_lo;
            ps0 &= ~31ull;
            if (ps0 < a.valid_lo) ps0 = (a.valid_lo + 31) & ~31ull;
            for (uint64_t b = ps0; b < seg_start; b += 32) insert_batch(S, a, b, seg_start, a.valid_lo);
        }
        __syncthreads();

        for (uint32_t c = c0; c < c1; ++c) {
            const uint64_t cbase = a.org + a.in_off[c];
            const uint64_t cend = a.org + a.in_off[c + 1];
            const uint32_t n = (uint32_t)(cend - cbase);
            const uint32_t nsteps = (n + kStep - 1) / kStep;
            const uint64_t lo = a.cross ? a.valid_lo : cbase;
            ParseState ps = {0, 0, 0, 0, 0, 0};

            for (uint32_t k = 0; k < nsteps + 2; ++k) {
                if (wid == kWarpInsert) {
                    if (k < nsteps) {
                        const uint64_t sb = cbase + (uint64_t)k * kStep;
                        const uint64_t se = sb + kStep < cend ? sb + kStep : cend;
                        for (uint64_t b = sb; b < se; b += 32) insert_batch(S, a, b, se, lo);
                    }
                } else if (wid == kWarpParse) {
                    if (k >= 1) {
                        const uint64_t av = (uint64_t)(k - 1) * kStep;
                        const uint32_t avail = av < n ? (uint32_t)av : n;
                        while (ps.ppos < avail && (ps.ppos + 32 < avail || avail == n)) {
                            parse_batch(S, a, cfg, ps, c, cbase, n, ps.ppos);
                            ps.ppos += 32;
                        }
                    }
                } else {
                    if (k >= 1 && k <= nsteps) {
                        const uint32_t j = k - 1;
                        const uint32_t i = (wid - 2) * 32 + lane;
                        const uint64_t q = (uint64_t)j * kStep + i;
                        if (q < n) S.res[(j % 3) * kStep + i] = search_position(S, a, cfg, cbase + q, cend, lo);
                    }
                }
                __syncthreads();
            }
            if (wid == kWarpParse) {
                close_block(S, a, ps, c, n);  // the final (possibly empty) block of the chunk
                if (lane == 0) a.chunk_nblk[c] = ps.blk < a.max_bpc ? ps.blk : a.max_bpc;
            }
            __syncthreads();
        }
    }
}

}  // namespace

int zs_launch_lz77(zs_ctx* ctx, const zs_deflate_plan& p) {
    static bool attr_set = false;
    if (!attr_set) {
        ZS_CUDA_TRY(ctx, cudaFuncSetAttribute(lz77_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)sizeof(Smem)));
        attr_set = true;
    }
    LzArgs a;
    const uintptr_t first = reinterpret_cast<uintptr_t>(p.d_in) - p.history;
    const uintptr_t adj = first & 15u;
    a.buf = reinterpret_cast<const uint8_t*>(first - adj);
    a.valid_lo = adj;
    a.org = adj + p.history;
    a.data_end = a.org + p.in_len;
    a.safe_end = (a.data_end + 7) & ~7ull;
    a.in_off = p.d_in_off;
    a.n_chunks = p.n_chunks;
    a.max_bpc = p.max_bpc;
    a.level = p.level;
    a.cross = (p.mode == ZS_MODE_STITCHED || (p.flags & ZS_FLAG_PRIME)) ? 1 : 0;
    // segments: runs of consecutive chunks sharing one set of hash tables.  With cross-chunk
    // matching a segment pays one 32 KiB insert-only priming pass; ~4 segments per SM keep the
    // dynamic scheduler balanced, and small batches fall back to one chunk per segment.
    uint32_t seg_chunks = 1;
    if (a.cross) {
        seg_chunks = p.n_chunks / (4u * (uint32_t)ctx->sm_count);
        if (seg_chunks < 1) seg_chunks = 1;
    }
    a.seg_chunks = seg_chunks;
    a.n_seg = (p.n_chunks + seg_chunks - 1) / seg_chunks;
    a.sym = p.d_sym;
    a.chunk_nblk = p.d_chunk_nblk;
    a.blk_desc = p.d_blk_desc;
    a.blk_freq = p.d_blk_freq;
    a.seg_counter = p.d_seg_counter;
    ZS_CUDA_TRY(ctx, cudaMemsetAsync(p.d_seg_counter, 0, sizeof(uint32_t), ctx->stream));
    unsigned grid = a.n_seg < (unsigned)ctx->sm_count ? a.n_seg : (unsigned)ctx->sm_count;
    if (grid == 0) return ZS_OK;
    lz77_kernel<<<grid, kThreads, sizeof(Smem), ctx->stream>>>(a);
    ZS_LAUNCH_CHECK(ctx, "lz77_kernel");
    return ZS_OK;
}
