// zs_lz77.cu -- K1/K2: LZ77 match finding (greedy levels 1-3, lazy levels 4-9) and parse for
// independent or dictionary-primed chunks.  (K3, the symbol histogram, is histogram_kernel in
// zs_huff.cu.)
//
// Replaces, for whole chunks, the reference's per-byte loop: INSERT_STRING / hash chains
// (src/mod/deflate/deflate.ts:109-141), longest_match (:1053-1115), deflate_fast (:1281-1350),
// deflate_slow incl. the TOO_FAR rule (:1352-1448), _tr_tally_lit/_tr_tally_dist
// (src/mod/deflate/utils.ts:55-81), with CONFIGURATION_TABLE (deflate.ts:86-103) giving
// max_chain / nice_length / max_lazy per level.  The bit stream differs from the reference's (the
// north star allows it); what must hold is: the symbols decode to the input, and the compressed
// size stays within 3 % of the reference at the same level (tests/test_deflate_gpu.py).
//
// Work decomposition (B200: 148 SMs, 227 KB shared memory per CTA):
//   * one persistent CTA per SM pulls *segments* (runs of consecutive chunks) from an atomic
//     counter.  Everything the per-byte loop touches lives in shared memory: a 64 KiB ring of the
//     window (staged from HBM once, 16 bytes per lane, coalesced) and the 2 x 64 KiB hash-chain
//     tables head[32768] / prev[32768] (16-bit positions like the reference's).  Tables and ring
//     are carried from chunk to chunk inside a segment, so dictionary priming costs one 32 KiB
//     insert-only pass per segment, not per chunk.
//   * inside the CTA the serial loop of the reference becomes a 5-stage software pipeline over
//     448-position steps, one __syncthreads per step.  Fourteen "wide" warps do everything that is
//     independent per position, two "thin" warps do the two inherently ordered updates:
//        stage 1  prep    (wide) : hash 32 positions, find equal hashes inside the batch with 15
//                                  ballots (nearest lower peer / last of its group);
//        stage 2  insert  (thin) : thread the chains through head/prev in position order -- a
//                                  single warp, ~25 instructions per 32 positions;
//        stage 3  search  (wide) : one lane per position walks that position's chain (<= max_chain
//                                  candidates, 8-byte compares in the shared-memory ring); every
//                                  position is searched speculatively, which is what makes the
//                                  chain walk 448-wide;
//        stage 4  resolve (wide) : applies the greedy / lazy rule per position and runs 5 rounds of
//                                  pointer doubling so that, for every possible entry lane of a
//                                  32-position batch, the visited positions and the exit are known;
//        stage 5  parse   (thin) : chains the batches (entry of batch b+1 = exit of batch b),
//                                  writes the packed symbols and cuts blocks every <= 16383
//                                  symbols (lit_bufsize - 1, deflate.ts:323,336).
//     The inserter never runs more than 2 steps ahead of a searcher, so restricting matches to
//     32768 - 2*448 bytes keeps every prev[] slot a searcher reads stable: no races, and the output
//     does not depend on scheduling (tables are reset per segment).
#include <cstdio>
#include <cstdlib>

#include "zs_common.cuh"

namespace {

#ifndef ZS_SEARCH_WARPS
#define ZS_SEARCH_WARPS 14
#endif
#ifndef ZS_HASH_BITS
#define ZS_HASH_BITS 15
#endif
constexpr int kSearchWarps = ZS_SEARCH_WARPS;
// The two thin roles get the highest warp ids: the SMSP arbiter favours high warp ids and these
// short critical-path warps must never wait behind the wide warps.
constexpr int kWarpParse = kSearchWarps;
constexpr int kWarpInsert = kSearchWarps + 1;
constexpr int kThreads = 32 * (kSearchWarps + 2);
constexpr int kStep = 32 * kSearchWarps;               // positions per pipeline step
constexpr unsigned kMaxDist = 32768u - 2u * kStep;      // see header comment
constexpr unsigned kSymLimit = 16383u;                  // symbols per block, deflate.ts:336
constexpr unsigned kTooFar = 4096u;                     // deflate/constants.ts (TOO_FAR)
constexpr unsigned kHashBits = ZS_HASH_BITS;
constexpr unsigned kRing = 65536;      // window ring: position p lives at ring[p & 0xffff]
constexpr unsigned kRingGuard = 288;   // ring[kRing + i] mirrors ring[i] so that 8-byte reads never wrap
// position-indexed rings between the stages (power-of-two sizes: cheap indexing)
constexpr unsigned pow2_at_least(unsigned v) { unsigned r = 1; while (r < v) r <<= 1; return r; }
constexpr unsigned kResRing = pow2_at_least(3 * kStep + 64);  // written by search, read by resolve and (two steps later) parse
constexpr unsigned kMjRing = pow2_at_least(2 * kStep + 64);   // written by resolve, read by parse one step later

struct LevelCfg { int lazy_fn, good, lazy, nice, chain; };
// CONFIGURATION_TABLE, deflate.ts:86-103
__constant__ LevelCfg c_levels[10] = {
    {0, 0, 0, 0, 0},      {0, 4, 4, 8, 4},      {0, 4, 5, 16, 8},      {0, 4, 6, 32, 32},
    {1, 4, 4, 16, 16},    {1, 8, 16, 32, 32},   {1, 8, 16, 128, 128},  {1, 8, 32, 128, 256},
    {1, 32, 128, 258, 1024}, {1, 32, 258, 258, 4096}};

// prep word: hash | nearest lower peer lane << 16 | has peer << 21 | last of its group << 22 | valid << 23
constexpr uint32_t PREP_HAS_PRED = 1u << 21, PREP_LAST = 1u << 22, PREP_VALID = 1u << 23;

struct Smem {
    uint8_t ring[kRing + kRingGuard];
    uint16_t head[1 << kHashBits];
    uint16_t prev[32768];
    uint32_t prep[2 * kStep];  // stage 1 -> 2, double buffered by step parity
    uint32_t res[kResRing];    // stage 3 -> 4,5: dist (15) | match len << 15 (9, 0 = none) | literal << 24
    uint2 mj[kMjRing];         // stage 4 -> 5: (visited mask, exit | is_match << 16) of a parse entering the batch at this lane
    uint32_t seg;              // segment being processed
};
static_assert(sizeof(Smem) <= 227 * 1024, "shared memory budget");

struct LzArgs {
    const uint8_t* buf;        // 16-byte aligned; position 0 of all indices below
    uint64_t org;              // index of the call's first input byte (d_in[0]) inside buf
    uint64_t valid_lo;         // first readable index (org - history)
    uint64_t data_end;         // one past the last input byte
    const uint64_t* in_off;    // [n_chunks+1] relative to org
    uint32_t n_chunks, seg_chunks, n_seg, max_bpc;
    int level;
    int cross;                 // 1: matches may reach before the chunk start (PRIME / STITCHED)
    uint32_t* sym;
    uint32_t* chunk_nblk;
    uint32_t* blk_desc;
    uint32_t* seg_counter;
    int debug;                 // ZS_LZ_PROF builds: chain override in bits 8..
};

#ifdef ZS_LZ_PROF
// cycle counters: [0] thin insert, [1] thin parse, [2] wide warps (summed), [3] whole step incl.
// barrier (warp 0), [4] steps
__device__ unsigned long long g_prof[16];
#define PROF_T(i) if (lane == 0) atomicAdd(&g_prof[i], (unsigned long long)(clock64() - t_begin))
#else
#define PROF_T(i)
#endif

__device__ __forceinline__ unsigned res_slot(uint32_t q) { return q & (kResRing - 1u); }
__device__ __forceinline__ unsigned mj_slot(uint32_t q) { return q & (kMjRing - 1u); }

// ---- window staging -------------------------------------------------------------------------------
// The match finder reads the window byte-granular and at random (one candidate per lane), which
// through L1 costs one 128-byte line per lane per load and saturates the L1 wavefront queue.  The
// window is therefore staged once, coalesced (16 bytes per lane), into a 64 KiB shared-memory ring
// that always holds the last 32 KiB plus the look-ahead; all compares are shared-memory loads.
__device__ __forceinline__ uint64_t win64(const Smem& S, uint64_t pos) {
    const unsigned idx = (unsigned)pos & (kRing - 1u);
    const unsigned al = idx & ~7u;
    const unsigned sh = (idx & 7u) * 8u;
    const uint64_t lo = *reinterpret_cast<const uint64_t*>(S.ring + al);
    const uint64_t hi = *reinterpret_cast<const uint64_t*>(S.ring + al + 8);
    return sh ? ((lo >> sh) | (hi << (64u - sh))) : lo;
}
__device__ __forceinline__ uint32_t win32(const Smem& S, uint64_t pos) {
    const unsigned idx = (unsigned)pos & (kRing - 1u);
    const unsigned al = idx & ~3u;
    const uint32_t lo = *reinterpret_cast<const uint32_t*>(S.ring + al);
    const uint32_t hi = *reinterpret_cast<const uint32_t*>(S.ring + al + 4);
    return __funnelshift_r(lo, hi, (idx & 3u) * 8u);
}

// Copy input [from, to) (absolute indices, multiples of 16) into the ring; `tid`/`nthr` describe
// the cooperating threads.  Bytes past the end of the input are staged as zeros.
__device__ __forceinline__ void stage_window(Smem& S, const LzArgs& a, uint64_t from, uint64_t to, unsigned tid, unsigned nthr) {
    const uint64_t safe16 = (a.data_end + 15) & ~15ull;
    for (uint64_t pos = from + 16ull * tid; pos < to; pos += 16ull * nthr) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (pos < safe16) v = __ldg(reinterpret_cast<const uint4*>(a.buf + pos));
        const unsigned idx = (unsigned)pos & (kRing - 1u);
        *reinterpret_cast<uint4*>(S.ring + idx) = v;
        if (idx < kRingGuard) *reinterpret_cast<uint4*>(S.ring + kRing + idx) = v;
    }
}

__device__ __forceinline__ unsigned hash3(uint32_t w) { return ((w & 0xffffffu) * 0x9E3779B1u) >> (32 - kHashBits); }

// ---- stage 1: prep (wide) -----------------------------------------------------------------------
// 32 consecutive positions starting at b; positions >= limit (end of the range being inserted) or
// without three readable bytes are not inserted.
__device__ __forceinline__ uint32_t prep_batch(const Smem& S, const LzArgs& a, uint64_t b, uint64_t limit) {
    const unsigned lane = zs_lane();
    const uint64_t p = b + lane;
    const bool valid = p < limit && p + 2 < a.data_end;
    unsigned h = 0;
    if (valid) h = hash3(win32(S, p));
    // lanes with the same hash: kHashBits independent ballots
    unsigned peers = __ballot_sync(ZS_FULL_MASK, valid);
#pragma unroll
    for (int bit = 0; bit < (int)kHashBits; ++bit) {
        const bool one = (h >> bit) & 1u;
        const unsigned bal = __ballot_sync(ZS_FULL_MASK, one);
        peers &= one ? bal : ~bal;
    }
    if (!valid) return 0u;
    const unsigned lower = peers & zs_lanemask_lt();
    uint32_t w = h | PREP_VALID;
    if (lower) w |= PREP_HAS_PRED | ((31u - __clz(lower)) << 16);
    if ((peers >> lane) == 1u) w |= PREP_LAST;
    return w;
}

// ---- stage 2: insert (thin) -----------------------------------------------------------------------
// Chains are kept in position order: the predecessor of p is the nearest earlier position with the
// same hash -- inside the batch (from prep) or head[h]; "none" is encoded as a self link.
// Only the head[] accesses are ordered (read old head, then publish the batch's last position of
// every hash); they do not depend on any loaded value, so a whole step's reads and writes are
// issued back to back and the loaded heads are consumed afterwards (link_batch).
__device__ __forceinline__ unsigned head_exchange(Smem& S, uint64_t b, uint32_t w) {
    const unsigned h = w & 0x7fffu;
    unsigned old = 0;
    if ((w & PREP_VALID) && !(w & PREP_HAS_PRED)) old = S.head[h];
    if (w & PREP_LAST) S.head[h] = (uint16_t)(b + zs_lane());
    __syncwarp();  // orders this batch's head[] stores before the next batch's loads
    return old;
}
__device__ __forceinline__ void link_batch(Smem& S, uint64_t b, uint64_t lo, uint32_t w, unsigned old) {
    if (w & PREP_VALID) {
        const uint64_t p = b + zs_lane();
        uint64_t pred = p;
        if (w & PREP_HAS_PRED) {
            pred = b + ((w >> 16) & 31u);
        } else {
            const unsigned delta = ((unsigned)p - old) & 0xffffu;
            if (delta != 0 && delta <= kMaxDist && p >= lo + delta) pred = p - delta;
        }
        S.prev[p & 32767u] = (uint16_t)pred;
    }
}

// ---- stage 3: search (wide), one lane per position -----------------------------------------------
// All arithmetic is on 16-bit ring indices and 32-bit distances (the ring is a multiple of the
// 32 KiB prev[] period, so prev[] is indexed by the ring index too).
__device__ __forceinline__ uint32_t ring32(const Smem& S, unsigned idx) {
    idx &= kRing - 1u;
    const unsigned al = idx & ~3u;
    const uint32_t lo = *reinterpret_cast<const uint32_t*>(S.ring + al);
    const uint32_t hi = *reinterpret_cast<const uint32_t*>(S.ring + al + 4);
    return __funnelshift_r(lo, hi, (idx & 3u) * 8u);
}
__device__ __forceinline__ uint64_t ring64(const Smem& S, unsigned idx) {
    idx &= kRing - 1u;
    const unsigned al = idx & ~3u, sh = (idx & 3u) * 8u;
    const uint32_t w0 = *reinterpret_cast<const uint32_t*>(S.ring + al);
    const uint32_t w1 = *reinterpret_cast<const uint32_t*>(S.ring + al + 4);
    const uint32_t w2 = *reinterpret_cast<const uint32_t*>(S.ring + al + 8);
    return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
}
__device__ __forceinline__ unsigned first_diff_byte(uint32_t x) { return (unsigned)(__ffs((int)x) - 1) >> 3; }

__device__ __forceinline__ uint32_t search_position(const Smem& S, const LevelCfg& cfg, uint64_t p, uint64_t chunk_end,
                                                    uint64_t lo) {
    const uint64_t room = chunk_end - p;
    const unsigned max_len = room < 258 ? (unsigned)room : 258u;
    const unsigned pi = (unsigned)p & (kRing - 1u);
    const uint32_t pw0 = ring32(S, pi), pw1 = ring32(S, pi + 4);
    const uint32_t lit = (pw0 & 0xffu) << 24;
    if (max_len < 3) return lit;
    const unsigned nice = (unsigned)cfg.nice < max_len ? (unsigned)cfg.nice : max_len;
    const uint64_t back = p - lo;
    const unsigned max_back = back < kMaxDist ? (unsigned)back : kMaxDist;
    unsigned best_len = 2, best_dist = 0;
    unsigned ci = pi, dist = 0;
    for (int chain = cfg.chain; chain > 0; --chain) {
        const unsigned delta = (ci - S.prev[ci & 32767u]) & 0xffffu;
        if (delta == 0) break;
        dist += delta;
        if (dist > max_back) break;
        ci = (ci - delta) & (kRing - 1u);
        uint32_t x = ring32(S, ci) ^ pw0;
        if ((x & 0xffffffu) != 0) continue;  // hash collision
        unsigned len;
        if (x) {
            len = 3;
        } else {
            x = ring32(S, ci + 4) ^ pw1;
            if (x) {
                len = 4 + first_diff_byte(x);
            } else {
                len = 8;
                while (len < max_len) {
                    const uint64_t y = ring64(S, ci + len) ^ ring64(S, pi + len);
                    if (y) {
                        const uint32_t yl = (uint32_t)y;
                        len += yl ? first_diff_byte(yl) : 4 + first_diff_byte((uint32_t)(y >> 32));
                        break;
                    }
                    len += 8;
                }
            }
        }
        if (len > max_len) len = max_len;
        if (len > best_len) {
            best_len = len;
            best_dist = dist;
            if (len >= nice) break;
        }
    }
    if (best_len < 3) return lit;
    if (cfg.lazy_fn && best_len == 3 && best_dist > kTooFar) return lit;  // deflate.ts:1381-1387
    return lit | (best_len << 15) | best_dist;
}

// ---- stage 4: resolve (wide) -----------------------------------------------------------------------
// Batch of 32 positions starting at q0 (chunk-relative, n bytes in the chunk).  Needs the search
// result of position q0+32 (or q0+32 >= n).
__device__ __forceinline__ void resolve_batch(Smem& S, const LevelCfg& cfg, uint32_t n, uint32_t q0) {
    const unsigned lane = zs_lane();
    const uint32_t q = q0 + lane;
    const uint32_t r = q < n ? S.res[res_slot(q)] : 0u;
    const unsigned L = (r >> 15) & 0x1ffu;
    unsigned Ln = __shfl_down_sync(ZS_FULL_MASK, L, 1);
    if (lane == 31) Ln = (q + 1 < n) ? ((S.res[res_slot(q + 1)] >> 15) & 0x1ffu) : 0u;
    // deflate_slow's lazy evaluation (deflate.ts:1372-1426): the match at q is dropped for a
    // literal when the match at q+1 is strictly longer and L < max_lazy
    const bool deferred = cfg.lazy_fn && L >= 3 && L < (unsigned)cfg.lazy && Ln > L;
    const bool is_match = L >= 3 && !deferred;
    unsigned J = lane + (is_match ? L : 1u);
    unsigned M = 1u << lane;
#pragma unroll
    for (int round = 0; round < 5; ++round) {
        const unsigned src = J < 32 ? J : lane;
        const unsigned Mj = __shfl_sync(ZS_FULL_MASK, M, src);
        const unsigned Jj = __shfl_sync(ZS_FULL_MASK, J, src);
        if (J < 32) { M |= Mj; J = Jj; }
    }
    if (q < n) S.mj[mj_slot(q)] = make_uint2(M, J | (is_match ? 0x10000u : 0u));
}

// ---- stage 5: parse (thin) ---------------------------------------------------------------------------
struct ParseState {
    uint32_t ppos;       // next batch start, relative to the chunk
    uint32_t skip;       // leading positions of that batch already covered by an emitted match
    uint32_t nsym;       // symbols emitted in the chunk
    uint32_t blk;        // blocks closed in the chunk
    uint32_t blk_sym0;   // first symbol of the open block
    uint32_t blk_pos0;   // first input position of the open block
};

__device__ __forceinline__ void close_block(const LzArgs& a, ParseState& ps, uint32_t chunk, uint32_t pos_end) {
    if (ps.blk < a.max_bpc && zs_lane() == 0) {  // max_bpc is sized so that the test always holds
        uint32_t* d = a.blk_desc + ((uint64_t)chunk * a.max_bpc + ps.blk) * 4;
        d[0] = ps.blk_sym0;
        d[1] = ps.nsym - ps.blk_sym0;
        d[2] = ps.blk_pos0;
        d[3] = pos_end - ps.blk_pos0;
    }
    ps.blk++;
    ps.blk_sym0 = ps.nsym;
    ps.blk_pos0 = pos_end;
}

__device__ __forceinline__ void parse_batch(const Smem& S, const LzArgs& a, ParseState& ps, uint32_t chunk,
                                            uint64_t sym_base, uint32_t n, uint32_t q0) {
    const unsigned lane = zs_lane();
    if (ps.skip >= 32) { ps.skip -= 32; return; }
    const uint2 e = S.mj[mj_slot(q0 + ps.skip)];  // same address in all lanes: a broadcast
    unsigned visited = e.x;
    if (n - q0 < 32) visited &= (1u << (n - q0)) - 1u;
    const unsigned nv = __popc(visited);
    if (ps.nsym - ps.blk_sym0 + nv > kSymLimit) {
        // the open block ends before this batch's first emitted position
        close_block(a, ps, chunk, q0 + (__ffs(visited) - 1u));
    }
    if ((visited >> lane) & 1u) {
        const uint32_t r = S.res[res_slot(q0 + lane)];
        const bool is_match = (S.mj[mj_slot(q0 + lane)].y >> 16) & 1u;
        // packed symbol: distance << 16 | length, or the literal byte (distance 0)
        a.sym[sym_base + ps.nsym + __popc(visited & zs_lanemask_lt())] =
            is_match ? (((r & 0x7fffu) << 16) | ((r >> 15) & 0x1ffu)) : (r >> 24);
    }
    ps.nsym += nv;
    ps.skip = (n - q0 < 32) ? 0u : (e.y & 0xffffu) - 32u;
}

// The same for a group of up to kParseGroup batches, arranged for instruction-level parallelism: the
// only truly serial chain is skip -> mj lookup -> next skip (one shared-memory load per batch);
// the symbol stores of all batches are independent once that chain is known.
constexpr int kParseGroup = 16;
__device__ __forceinline__ void parse_group(const Smem& S, const LzArgs& a, ParseState& ps, uint32_t chunk,
                                            uint64_t sym_base, uint32_t n, uint32_t qb, int nb, long long t_begin) {
    const unsigned lane = zs_lane();
    unsigned vis[kParseGroup];
    uint32_t base[kParseGroup];
    uint2 mine[kParseGroup];   // (visited, exit | is_match << 16) for a parse entering at this lane
#pragma unroll
    for (int u = 0; u < kParseGroup; ++u) {
        const uint32_t q = qb + 32u * u + lane;
        mine[u] = (u < nb && q < n) ? S.mj[mj_slot(q)] : make_uint2(0u, 32u);
    }
    uint32_t nsym = ps.nsym, skip = ps.skip;
#pragma unroll
    for (int u = 0; u < kParseGroup; ++u) {
        vis[u] = 0;
        base[u] = nsym;
        if (u < nb) {
            const uint32_t q0 = qb + 32u * u;
            if (skip >= 32) {
                skip -= 32;
            } else {
                // the serial chain: one register shuffle per batch
                unsigned v = __shfl_sync(ZS_FULL_MASK, mine[u].x, skip);
                const unsigned ex = __shfl_sync(ZS_FULL_MASK, mine[u].y, skip) & 0xffffu;
                if (n - q0 < 32) v &= (1u << (n - q0)) - 1u;
                vis[u] = v;
                nsym += __popc(v);
                skip = (n - q0 < 32) ? 0u : ex - 32u;
            }
        }
    }
#ifdef ZS_LZ_PROF
    if (lane == 0) atomicAdd(&g_prof[11], (unsigned long long)(clock64() - t_begin));
#endif
    if (nsym - ps.blk_sym0 > kSymLimit) {
        // a block boundary falls inside this group: take the batch-by-batch path
        for (int u = 0; u < nb; ++u) parse_batch(S, a, ps, chunk, sym_base, n, qb + 32u * u);
        return;
    }
#pragma unroll
    for (int u = 0; u < kParseGroup; ++u) {
        if ((vis[u] >> lane) & 1u) {
            const uint32_t q = qb + 32u * u + lane;
            const uint32_t r = S.res[res_slot(q)];
            const bool is_match = (mine[u].y >> 16) & 1u;
            a.sym[sym_base + base[u] + __popc(vis[u] & zs_lanemask_lt())] =
                is_match ? (((r & 0x7fffu) << 16) | ((r >> 15) & 0x1ffu)) : (r >> 24);
        }
    }
    ps.nsym = nsym;
    ps.skip = skip;
}

// Searched frontier (exclusive, chunk-relative) at the start of pipeline iteration k: the search
// of step j runs in iteration j + 2.
__device__ __forceinline__ uint32_t searched_at(uint32_t k, uint32_t n) {
    if (k < 3) return 0;
    const uint64_t f = (uint64_t)(k - 2) * kStep;
    return f < n ? (uint32_t)f : n;
}
// Resolved frontier after iteration k: batches whose successor position has been searched.
__device__ __forceinline__ uint32_t resolved_after(uint32_t k, uint32_t n) {
    const uint32_t f = searched_at(k, n);
    if (f == n) return n;
    return f >= 32 ? f - 32 : 0;
}

__global__ void __launch_bounds__(kThreads, 1) lz77_kernel(LzArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const unsigned wid = threadIdx.x >> 5, lane = zs_lane();
    LevelCfg cfg = c_levels[a.level];
#ifdef ZS_LZ_PROF
    if (a.debug >> 8) cfg.chain = a.debug >> 8;
#endif

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
        }
        // Range 0 of a segment is the dictionary (deflateSetDictionary, deflate.ts:367-424): the
        // <= 32 KiB before the first chunk go through prep + insert only.
        const uint64_t seg_start = a.org + a.in_off[c0];
        uint64_t prime0 = seg_start;
        if (a.cross) {
            prime0 = seg_start > a.valid_lo + 32768 ? seg_start - 32768 : a.valid_lo;
            prime0 = (prime0 + 31) & ~31ull;
            if (prime0 > seg_start) prime0 = seg_start;
        }
        uint64_t staged_end = prime0 & ~15ull;  // [staged_end - 64 KiB, staged_end) is in the ring
        __syncthreads();

        for (int64_t ci = (int64_t)c0 - 1; ci < (int64_t)c1; ++ci) {
            const bool priming = ci < (int64_t)c0;
            const uint32_t c = priming ? c0 : (uint32_t)ci;
            const uint64_t cbase = priming ? prime0 : a.org + a.in_off[c];
            const uint64_t cend = priming ? seg_start : a.org + a.in_off[c + 1];
            if (priming && cbase == cend) continue;
            const uint32_t n = (uint32_t)(cend - cbase);
            const uint32_t nsteps = (n + kStep - 1) / kStep;
            const uint64_t lo = a.cross ? a.valid_lo : cbase;
            const uint64_t sym_base = cbase - a.org;
            const uint32_t n_iter = priming ? nsteps + 1 : nsteps + 4;
            ParseState ps = {0, 0, 0, 0, 0, 0};
            {   // look-ahead for the first two steps of the range (no-op when already staged)
                const uint64_t target = (cbase + 2ull * kStep + kRingGuard + 15) & ~15ull;
                if (target > staged_end) {
                    stage_window(S, a, staged_end, target, threadIdx.x, kThreads);
                    staged_end = target;
                }
                __syncthreads();
            }

            for (uint32_t k = 0; k < n_iter; ++k) {
#ifdef ZS_LZ_PROF
                const long long t_begin = clock64();
#else
                const long long t_begin = 0;
#endif
                if (wid == kWarpInsert) {
                    // Bytes the next iteration reads (prep of step k+1, look-ahead of the search of step
                    // k-1) replace positions 64 KiB older, which nobody reads any more.  The global
                    // loads are issued first and stored last so that their latency hides behind the
                    // table updates.
                    constexpr int kStageVec = (kStep + 16 + 511) / 512;   // 16-byte vectors per lane and step
                    uint4 sv[kStageVec];
                    uint64_t stage_from = staged_end, stage_to = staged_end;
                    if (k < nsteps) {
                        const uint64_t target = (cbase + (uint64_t)(k + 3) * kStep + kRingGuard + 15) & ~15ull;
                        if (target > staged_end) stage_to = target;
                    }
                    const uint64_t safe16 = (a.data_end + 15) & ~15ull;
#pragma unroll
                    for (int v = 0; v < kStageVec; ++v) {
                        const uint64_t pos = stage_from + 16ull * (lane + 32 * v);
                        sv[v] = make_uint4(0, 0, 0, 0);
                        if (pos < stage_to && pos < safe16) sv[v] = __ldg(reinterpret_cast<const uint4*>(a.buf + pos));
                    }
                    PROF_T(8);
                    if (k >= 1 && k - 1 < nsteps) {
                        const uint64_t sb = cbase + (uint64_t)(k - 1) * kStep;
                        const uint32_t* pw = S.prep + ((k - 1) & 1u) * kStep;
                        uint32_t w[kSearchWarps];
#pragma unroll
                        for (int u = 0; u < kSearchWarps; ++u) w[u] = pw[32 * u + lane];
                        PROF_T(9);
                        unsigned old[kSearchWarps];
#pragma unroll
                        for (int u = 0; u < kSearchWarps; ++u) old[u] = head_exchange(S, sb + 32u * u, w[u]);
#pragma unroll
                        for (int u = 0; u < kSearchWarps; ++u) link_batch(S, sb + 32u * u, lo, w[u], old[u]);
                    }
                    PROF_T(10);
#pragma unroll
                    for (int v = 0; v < kStageVec; ++v) {
                        const uint64_t pos = stage_from + 16ull * (lane + 32 * v);
                        if (pos < stage_to) {
                            const unsigned idx = (unsigned)pos & (kRing - 1u);
                            *reinterpret_cast<uint4*>(S.ring + idx) = sv[v];
                            if (idx < kRingGuard) *reinterpret_cast<uint4*>(S.ring + kRing + idx) = sv[v];
                        }
                    }
                    // anything beyond kStageVec vectors per lane (only after a short first range)
                    if (stage_to > stage_from + 512ull * kStageVec) stage_window(S, a, stage_from + 512ull * kStageVec, stage_to, lane, 32);
                } else if (wid == kWarpParse) {
                    if (!priming && k >= 4) {
                        const uint32_t upto = resolved_after(k - 1, n);
                        while (ps.ppos < upto) {
                            int nb = (int)((upto - ps.ppos + 31) / 32);
                            if (nb > kParseGroup) nb = kParseGroup;
                            parse_group(S, a, ps, c, sym_base, n, ps.ppos, nb, t_begin);
                            ps.ppos += 32u * nb;
                        }
                    }
                } else {
                    // stage 1: prep of step k
                    if (k < nsteps) {
                        const uint64_t sb = cbase + (uint64_t)k * kStep;
                        const uint64_t se = sb + kStep < cend ? sb + kStep : cend;
                        S.prep[(k & 1u) * kStep + wid * 32 + lane] = prep_batch(S, a, sb + 32u * wid, se);
                    }
                    if (!priming) {
                        // stage 3: search of step k-2
                        if (k >= 2 && k - 2 < nsteps) {
                            const uint32_t q = (k - 2) * kStep + wid * 32 + lane;
                            if (q < n) S.res[res_slot(q)] = search_position(S, cfg, cbase + q, cend, lo);
                        }
                        // stage 4: resolve the batches whose successor was searched before this iteration
                        if (k >= 3) {
                            const uint32_t from = resolved_after(k - 1, n), upto = resolved_after(k, n);
                            for (uint32_t q0 = from + 32u * wid; q0 < upto; q0 += 32u * kSearchWarps)
                                resolve_batch(S, cfg, n, q0);
                        }
                    }
                }
#ifdef ZS_LZ_PROF
                {
                    const long long t_work = clock64() - t_begin;
                    __syncwarp();
                    if (lane == 0) {
                        const int role = wid == kWarpInsert ? 0 : wid == kWarpParse ? 1 : 2;
                        atomicAdd(&g_prof[role], (unsigned long long)t_work);
                        if (wid == 0) atomicAdd(&g_prof[4], 1ull);
                    }
                }
#endif
                if (k < nsteps) {
                    const uint64_t target = (cbase + (uint64_t)(k + 3) * kStep + kRingGuard + 15) & ~15ull;
                    if (target > staged_end) staged_end = target;
                }
                __syncthreads();
#ifdef ZS_LZ_PROF
                if (threadIdx.x == 0) atomicAdd(&g_prof[3], (unsigned long long)(clock64() - t_begin));
#endif
            }
            if (!priming && wid == kWarpParse) {
                close_block(a, ps, c, n);  // the final (possibly empty) block of the chunk
                if (lane == 0) a.chunk_nblk[c] = ps.blk < a.max_bpc ? ps.blk : a.max_bpc;
            }
        }
        __syncthreads();
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
        if (seg_chunks > 16) seg_chunks = 16;  // fixed for large batches: slicing a batch at multiples of 16 chunks
                                               // (the pipelined host path) then leaves the output unchanged
    }
    if (a.cross && p.seg_hint) seg_chunks = p.seg_hint;
    a.seg_chunks = seg_chunks;
    a.n_seg = (p.n_chunks + seg_chunks - 1) / seg_chunks;
    a.sym = p.d_sym;
    a.chunk_nblk = p.d_chunk_nblk;
    a.blk_desc = p.d_blk_desc;
    a.seg_counter = p.d_seg_counter;
    a.debug = 0;
#ifdef ZS_LZ_PROF
    if (getenv("ZS_LZ_DEBUG")) a.debug = atoi(getenv("ZS_LZ_DEBUG"));
#endif
    ZS_CUDA_TRY(ctx, cudaMemsetAsync(p.d_seg_counter, 0, sizeof(uint32_t), ctx->stream));
    unsigned grid = a.n_seg < (unsigned)ctx->sm_count ? a.n_seg : (unsigned)ctx->sm_count;
    if (grid == 0) return ZS_OK;
#ifdef ZS_LZ_PROF
    unsigned long long zero[16] = {0};
    cudaMemcpyToSymbol(g_prof, zero, sizeof(zero));
#endif
    ZS_KERNEL(ctx, "lz77_kernel", lz77_kernel<<<grid, kThreads, sizeof(Smem), ctx->stream>>>(a));
#ifdef ZS_LZ_PROF
    {
        unsigned long long pr[16];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpyFromSymbol(pr, g_prof, sizeof(pr));
        double st = pr[4] ? (double)pr[4] : 1.0;
        fprintf(stderr, "[lz77 prof] level %d steps %llu  cycles/step: insert %.0f parse %.0f wide(avg warp) %.0f step %.0f | insert: ldg-issued %.0f prep-loaded %.0f inserted %.0f | parse chain done %.0f\n",
                p.level, pr[4], pr[0] / st, pr[1] / st, pr[2] / st / kSearchWarps, pr[3] / st, pr[8] / st, pr[9] / st, pr[10] / st, pr[11] / st);
    }
#endif
    return ZS_OK;
}
