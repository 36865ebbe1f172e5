/*
 * lzmodel.c -- CPU model of lz77_kernel's matching policy (zlib-streams-ts_b200/csrc/zs_lz77.cu), for
 * exploring policy changes without a GPU: it restates what the kernel decides -- every position
 * inserted (14-bit multiplicative hash of 3 bytes, position-ordered chains with 16-bit links), every
 * data position searched with search_position's rules, resolve_one's greedy / lazy choice, blocks cut
 * at 16383 symbols and at chunk ends -- and reports
 *   - the compressed size (per block: the cheapest of stored / fixed / dynamic Huffman, with its own
 *     length-limited Huffman construction: within ~0.1 % of the engine's encoder),
 *   - the work of the chain walk as the SIMT hardware sees it: candidates visited (lane iterations)
 *     and, per aligned batch of 32 positions, the longest walk (what the warp waits for).
 * Not part of the product, not the oracle: a design tool.  Build: gcc -O2 -o lzmodel lzmodel.c
 *
 * usage: lzmodel <file> <level> <chunk bytes> <chunks per segment> [key=value ...]
 *   stop_active=K stop_after=M   a batch stops walking once at most K lanes are still walking and
 *                                at least M candidates have been visited (0 = off)
 *   dense_hop=H interior=C       the two constants of the lazy levels' work bounds (8, 16)
 *   work_shift=S                 lazy levels: every 8-byte compare round also takes 1/2^S off the chain budget (-1 = off)
 *   skip_lag=D skip_cap=N        greedy levels: a candidate at distance >= D that the parse did not visit and that lies
 *                                inside a match longer than max_insert (deflate_fast would not have inserted it,
 *                                deflate.ts:1310-1322) is passed over without using up the chain budget; at most N per position
 *   too_far=D too_far4=D         greedy levels: 3-byte (4-byte) matches at a distance above D become literals
 *   chain=N nice=N               override the level's max_chain / nice_length
 *   probe=K                      greedy levels, "parse before search": every position is PROBED with K candidates only, a
 *                                greedy parse over the probe results marks the positions it visits, only those get the full
 *                                chain walk, and the final parse uses the full result where there is one and the probe
 *                                result elsewhere (a longer match moves the parse onto positions that were only probed).
 *                                Reports the share of positions that are searched in full and the walk of the COMPACTED
 *                                batches (32 visited positions per warp).
 *   dump=path                    write the symbols (lit<<24 | len<<15 | dist) as little-endian u32
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { kStep = 960, kMaxDist = 32768 - 2 * kStep, kSymLimit = 16383, kTooFar = 4096, kHashBits = 14 };
typedef struct { int lazy_fn, good, lazy, nice, chain; } level_cfg;
static const level_cfg LEVELS[10] = {{0, 0, 0, 0, 0},      {0, 4, 4, 8, 4},      {0, 4, 5, 16, 8},     {0, 4, 6, 32, 32},
                                     {1, 4, 4, 16, 16},    {1, 8, 16, 32, 32},   {1, 8, 16, 128, 128}, {1, 8, 32, 128, 256},
                                     {1, 32, 128, 258, 1024}, {1, 32, 258, 258, 4096}};

static const uint8_t* buf;      /* the whole input, absolute positions; 300 readable bytes of padding behind it */
static size_t total;
static uint16_t head[1 << kHashBits], prev16[32768];
static int dense_hop = 8, interior_chain = 16, stop_active = 0, stop_after = 0, work_shift = -1;
static int too_far_greedy = 4096, too_far4 = 0;   /* greedy levels: a 3-byte (4-byte) match further back than this becomes a literal */
static int chain_override = 0, nice_override = 0;   /* chain= / nice=: replace the level's max_chain / nice_length */
static int probe_k = 0;                   /* probe=K: two-phase search of the greedy levels */
static int skip_lag = 0, skip_cap = 16;   /* greedy levels: candidates the reference would not have inserted are passed over */
static uint8_t* mark;                     /* 1 = deflate_fast would have put this position into its hash chains */

static inline uint32_t ld32(size_t p) { uint32_t v; memcpy(&v, buf + p, 4); return v; }
static inline unsigned hash3(uint32_t w) { return ((w & 0xffffffu) * 0x9E3779B1u) >> (32 - kHashBits); }

/* prep + insert + link for one position (the pipeline keeps the chains in position order) */
static void insert_pos(size_t abs, size_t range0, uint32_t pre) {
    if (abs + 2 >= total) return;
    const unsigned h = hash3(ld32(abs));
    const unsigned p16 = (unsigned)abs & 0xffffu;
    const unsigned old = head[h];
    const unsigned delta = (p16 - old) & 0xffffu;
    unsigned pred = p16;
    if (delta != 0 && delta <= kMaxDist && delta <= (abs - range0) + pre) pred = (p16 - delta) & 0xffffu;
    prev16[p16 & 32767u] = (uint16_t)pred;
    head[h] = (uint16_t)p16;
}

/* search_position: returns lit<<24 | len<<15 | dist; *visited = candidates looked at, cap = iteration cap */
static uint32_t search_pos(const level_cfg* cfg, int lazy, size_t abs, size_t range0, uint32_t pre, size_t cs, size_t ce,
                           int cross, unsigned cap, unsigned* visited, unsigned* rounds) {
    const uint32_t room = (uint32_t)(ce - abs);
    const unsigned max_len = room < 258u ? room : 258u;
    const uint32_t lit = (uint32_t)buf[abs] << 24;
    *visited = 0;
    *rounds = 0;
    if (max_len < 3) return lit;
    const unsigned nice = (unsigned)cfg->nice < max_len ? (unsigned)cfg->nice : max_len;
    const uint32_t back = cross ? (uint32_t)(abs - range0) + pre : (uint32_t)(abs - cs);
    const unsigned max_back = back < kMaxDist ? back : kMaxDist;
    unsigned best_len = 2, best_dist = 0, dist = 0, skipped = 0;
    unsigned ci = (unsigned)abs & 0xffffu;
    const uint32_t pw0 = ld32(abs), pw1 = ld32(abs + 4);
    for (int chain = cfg->chain; chain > 0; --chain) {
        if (cap && *visited >= cap) break;
        ++*visited;
        const unsigned delta = (ci - prev16[ci & 32767u]) & 0xffffu;
        if (delta == 0) break;
        dist += delta;
        if (dist > max_back) break;
        ci = (ci - delta) & 0xffffu;
        const size_t cand = abs - dist;
        if (skip_lag && dist >= (unsigned)skip_lag && !mark[cand] && skipped < (unsigned)skip_cap) {
            ++skipped;
            ++chain;   /* free: undoes the loop's decrement */
            continue;
        }
        uint32_t x = ld32(cand) ^ pw0;
        if ((x & 0xffffffu) != 0) continue;
        unsigned len;
        if (x) {
            len = 3;
        } else {
            x = ld32(cand + 4) ^ pw1;
            if (x) {
                len = 4 + (unsigned)(__builtin_ctz(x) >> 3);
            } else {
                if (lazy && best_len >= 8 && buf[cand + best_len] != buf[abs + best_len]) {
                    if (delta <= (unsigned)dense_hop) chain -= chain >> 2;
                    continue;
                }
                len = 8;
                while (len < max_len && buf[cand + len] == buf[abs + len]) len++;
                *rounds += (len - 8) / 8 + 1;   /* 8-byte rounds of the compare loop */
                if (lazy && work_shift >= 0) chain -= (int)(((len - 8) / 8 + 1) >> work_shift);   /* long compares are charged to the chain budget */
            }
        }
        if (len > max_len) len = max_len;
        if (len > best_len) {
            if (lazy && best_len < (unsigned)cfg->good && len >= (unsigned)cfg->good) {
                const int interior = cand > 0 && abs > 0 && buf[cand - 1] == buf[abs - 1];
                if (interior || delta <= (unsigned)dense_hop) chain >>= 2;
                if (interior && chain > interior_chain) chain = interior_chain;
            }
            best_len = len;
            best_dist = dist;
            if (len >= nice) break;
        } else if (lazy && len == best_len && delta <= (unsigned)dense_hop) {
            chain -= chain >> 2;
        }
    }
    if (best_len < 3) return lit;
    if (lazy && best_len == 3 && best_dist > kTooFar) return lit;
    if (!lazy && too_far_greedy && best_len == 3 && best_dist > (unsigned)too_far_greedy) return lit;
    if (!lazy && too_far4 && best_len == 4 && best_dist > (unsigned)too_far4) return lit;
    return lit | (best_len << 15) | best_dist;
}

/* ---- block cost ---------------------------------------------------------------------------------- */
static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_X[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_X[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static int len_code(unsigned len) { int c = 28; while (LEN_BASE[c] > len) c--; return c; }
static int dist_code(unsigned d) { int c = 29; while (DIST_BASE[c] > d) c--; return c; }

/* Huffman code lengths limited to `limit` bits: plain Huffman, frequencies flattened until it fits */
static void huff_lengths(const uint32_t* freq_in, int n, int limit, uint8_t* len) {
    uint64_t f[288];
    for (int i = 0; i < n; i++) f[i] = freq_in[i];
    for (;;) {
        int parent[2 * 288], alive[288], na = 0;
        uint64_t w[2 * 288];
        int nodes = n;
        for (int i = 0; i < n; i++) { w[i] = f[i]; if (f[i]) alive[na++] = i; }
        memset(len, 0, (size_t)n);
        if (na == 0) return;
        if (na == 1) { len[alive[0]] = 1; return; }
        int heap[2 * 288], hn = 0;
        for (int i = 0; i < na; i++) heap[hn++] = alive[i];
        while (hn > 1) {
            int a = -1, b = -1;   /* the two lightest: n is small, a linear scan is enough */
            for (int i = 0; i < hn; i++) if (a < 0 || w[heap[i]] < w[heap[a]]) a = i;
            int na_ = heap[a]; heap[a] = heap[--hn];
            for (int i = 0; i < hn; i++) if (b < 0 || w[heap[i]] < w[heap[b]]) b = i;
            int nb_ = heap[b]; heap[b] = heap[--hn];
            w[nodes] = w[na_] + w[nb_];
            parent[na_] = parent[nb_] = nodes;
            heap[hn++] = nodes++;
        }
        const int root = nodes - 1;
        int worst = 0;
        for (int i = 0; i < na; i++) {
            int d = 0;
            for (int v = alive[i]; v != root; v = parent[v]) d++;
            len[alive[i]] = (uint8_t)d;
            if (d > worst) worst = d;
        }
        if (worst <= limit) return;
        for (int i = 0; i < n; i++) if (f[i]) f[i] = (f[i] >> 1) + 1;
    }
}

static uint64_t dynamic_header_bits(const uint8_t* ll, int nl, const uint8_t* dl, int nd) {
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t seq[320];
    int n = 0;
    for (int i = 0; i < nl; i++) seq[n++] = ll[i];
    for (int i = 0; i < nd; i++) seq[n++] = dl[i];
    uint32_t bf[19] = {0};
    uint64_t extra = 0;
    /* the two trees are scanned separately by scan_tree; one pass over each */
    int start = 0;
    for (int part = 0; part < 2; part++) {
        const int end = part == 0 ? nl : n;
        int i = start;
        while (i < end) {
            int j = i;
            while (j < end && seq[j] == seq[i]) j++;
            int run = j - i;
            if (seq[i] == 0) {
                while (run >= 11) { int r = run > 138 ? 138 : run; bf[18]++; extra += 7; run -= r; }
                if (run >= 3) { bf[17]++; extra += 3; run = 0; }
                bf[0] += (uint32_t)run;
            } else {
                bf[seq[i]]++; run--;
                while (run >= 3) { int r = run > 6 ? 6 : run; bf[16]++; extra += 2; run -= r; }
                bf[seq[i]] += (uint32_t)run;
            }
            i = j;
        }
        start = nl;
    }
    uint8_t bl[19];
    huff_lengths(bf, 19, 7, bl);
    int hclen = 19;
    while (hclen > 4 && bl[order[hclen - 1]] == 0) hclen--;
    uint64_t bits = 14 + 3u * (unsigned)hclen + extra;
    for (int i = 0; i < 19; i++) bits += (uint64_t)bf[i] * bl[i];
    return bits;
}

static uint64_t block_bits(const uint32_t* sym, size_t ns, size_t in_bytes) {
    uint32_t lf[288] = {0}, df[30] = {0};
    uint64_t extra = 0;
    for (size_t i = 0; i < ns; i++) {
        const unsigned len = (sym[i] >> 15) & 0x1ffu;
        if (len) {
            const int lc = len_code(len), dc = dist_code(sym[i] & 0x7fffu);
            lf[257 + lc]++; df[dc]++;
            extra += LEN_X[lc] + DIST_X[dc];
        } else {
            lf[sym[i] >> 24]++;
        }
    }
    lf[256] = 1;
    uint8_t ll[288], dl[30];
    huff_lengths(lf, 286, 15, ll);
    huff_lengths(df, 30, 15, dl);
    uint64_t dyn = extra, fix = extra;
    for (int i = 0; i < 286; i++) { dyn += (uint64_t)lf[i] * ll[i]; fix += (uint64_t)lf[i] * (i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8); }
    for (int i = 0; i < 30; i++) { dyn += (uint64_t)df[i] * dl[i]; fix += (uint64_t)df[i] * 5; }
    int nl = 286, nd = 30;
    while (nl > 257 && ll[nl - 1] == 0) nl--;
    while (nd > 1 && dl[nd - 1] == 0) nd--;
    dyn += dynamic_header_bits(ll, nl, dl, nd);
    uint64_t best = dyn < fix ? dyn : fix;
    const uint64_t stored = 8ull * in_bytes + 32 + 7;   /* LEN/NLEN + alignment on average */
    if (stored < best) best = stored;
    return best + 3;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: lzmodel <file> <level> <chunk> <seg_chunks> [key=value ...]\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    fseek(f, 0, SEEK_END); total = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* data = (uint8_t*)calloc(total + 300, 1);
    if (fread(data, 1, total, f) != total) { perror("read"); return 1; }
    fclose(f);
    buf = data;
    const int level = atoi(argv[2]);
    const size_t chunk = (size_t)atol(argv[3]);
    const size_t seg_chunks = (size_t)atol(argv[4]);
    const char* dump = NULL;
    int cross = 1;
    for (int i = 5; i < argc; i++) {
        if (!strncmp(argv[i], "stop_active=", 12)) stop_active = atoi(argv[i] + 12);
        else if (!strncmp(argv[i], "stop_after=", 11)) stop_after = atoi(argv[i] + 11);
        else if (!strncmp(argv[i], "dense_hop=", 10)) dense_hop = atoi(argv[i] + 10);
        else if (!strncmp(argv[i], "interior=", 9)) interior_chain = atoi(argv[i] + 9);
        else if (!strncmp(argv[i], "cross=", 6)) cross = atoi(argv[i] + 6);
        else if (!strncmp(argv[i], "work_shift=", 11)) work_shift = atoi(argv[i] + 11);
        else if (!strncmp(argv[i], "skip_lag=", 9)) skip_lag = atoi(argv[i] + 9);
        else if (!strncmp(argv[i], "skip_cap=", 9)) skip_cap = atoi(argv[i] + 9);
        else if (!strncmp(argv[i], "dump=", 5)) dump = argv[i] + 5;
        else if (!strncmp(argv[i], "too_far=", 8)) too_far_greedy = atoi(argv[i] + 8);
        else if (!strncmp(argv[i], "too_far4=", 9)) too_far4 = atoi(argv[i] + 9);
        else if (!strncmp(argv[i], "chain=", 6)) chain_override = atoi(argv[i] + 6);
        else if (!strncmp(argv[i], "nice=", 5)) nice_override = atoi(argv[i] + 5);
        else if (!strncmp(argv[i], "probe=", 6)) probe_k = atoi(argv[i] + 6);
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
    }
    level_cfg cfg_copy = LEVELS[level];
    if (chain_override) cfg_copy.chain = chain_override;
    if (nice_override) cfg_copy.nice = nice_override;
    const level_cfg* cfg = &cfg_copy;
    const int lazy = cfg->lazy_fn;
    FILE* fd = dump ? fopen(dump, "wb") : NULL;

    uint32_t* res = (uint32_t*)malloc(sizeof(uint32_t) * (seg_chunks * chunk + 64));
    uint32_t* sym = (uint32_t*)malloc(sizeof(uint32_t) * (chunk + 64));
    if (skip_lag) mark = (uint8_t*)calloc(total + 300, 1);
    uint64_t bits = 0, nsym_total = 0, cand_total = 0, simt_total = 0, batches = 0, nblocks = 0, stopped = 0, rounds_total = 0, simt_heavy = 0, step_heavy_total = 0, steps = 0;
    uint64_t hist[16] = {0};   /* longest walk of a batch, log2 buckets */
    uint64_t probe_cand = 0, full_positions = 0, full_cand = 0, full_batches = 0, full_simt = 0, data_positions = 0;
    unsigned cw[32], ncw = 0;   /* walks of the visited positions waiting to fill a compacted batch */
    if (probe_k && lazy) { fprintf(stderr, "probe= applies to the greedy levels\n"); return 2; }

    for (size_t seg_start = 0; seg_start < total; seg_start += seg_chunks * chunk) {
        size_t seg_end = seg_start + seg_chunks * chunk;
        if (seg_end > total) seg_end = total;
        size_t prime0 = seg_start;
        if (cross) {
            prime0 = seg_start > 32768 ? seg_start - 32768 : 0;
            prime0 = (prime0 + 31) & ~(size_t)31;
            if (prime0 > seg_start) prime0 = seg_start;
        }
        const uint32_t pre = cross ? (prime0 < kMaxDist ? (uint32_t)prime0 : kMaxDist) : 0;
        memset(head, 0, sizeof head);
        memset(prev16, 0, sizeof prev16);
        const size_t n = seg_end - prime0, q_data = seg_start - prime0;
        /* insert + search, batch by batch (aligned to the range start like the kernel's batches) */
        unsigned step_heavy = 0;
        /* resolve + parse + blocks, incrementally: the parse trails the search like in the pipeline */
        size_t pp = seg_start, pcs = seg_start, pce = seg_start + chunk < seg_end ? seg_start + chunk : seg_end, pns = 0, pblk0 = seg_start;
        if (mark) memset(mark + prime0, 1, seg_start - prime0);   /* deflateSetDictionary inserts every position */
#define PARSE_UNTIL(limit_)                                                                                              \
        while (pp < (limit_)) {                                                                                          \
            const uint32_t r = res[pp - seg_start];                                                                      \
            const unsigned L = (r >> 15) & 0x1ffu;                                                                       \
            const unsigned Ln = pp + 1 < pce ? (res[pp + 1 - seg_start] >> 15) & 0x1ffu : 0;                             \
            const int deferred = lazy && L >= 3 && L < (unsigned)cfg->lazy && Ln > L;                                    \
            if (pns == kSymLimit) { bits += block_bits(sym, pns, pp - pblk0); if (fd) fwrite(sym, 4, pns, fd); nblocks++; nsym_total += pns; pns = 0; pblk0 = pp; } \
            if (L >= 3 && !deferred) {                                                                                   \
                sym[pns++] = r;                                                                                          \
                if (mark) { mark[pp] = 1; if (L <= (unsigned)cfg->lazy) memset(mark + pp, 1, L); }                       \
                pp += L;                                                                                                 \
            } else { sym[pns++] = r & 0xff000000u; if (mark) mark[pp] = 1; pp += 1; }                                    \
            if (pp >= pce) {                                                                                             \
                bits += block_bits(sym, pns, pce - pblk0); if (fd) fwrite(sym, 4, pns, fd); nblocks++; nsym_total += pns; \
                pns = 0; pcs = pce; pce = pcs + chunk < seg_end ? pcs + chunk : seg_end; pblk0 = pcs; pp = pcs;          \
                if (pcs >= seg_end) break;                                                                               \
            }                                                                                                            \
        }
        size_t pp1 = seg_start;   /* probe=: the parse over the probe results */
        for (size_t q0 = 0; q0 < n; q0 += 32) {
            if (skip_lag && prime0 + q0 > seg_start + (size_t)skip_lag) PARSE_UNTIL(prime0 + q0 - (size_t)skip_lag);
            unsigned walk[32], rnd[32];
            unsigned longest = 0;
            for (unsigned l = 0; l < 32 && q0 + l < n; l++) insert_pos(prime0 + q0 + l, prime0, pre);
            for (unsigned l = 0; l < 32; l++) {
                const size_t q = q0 + l;
                walk[l] = rnd[l] = 0;
                if (q < q_data || q >= n) continue;
                const size_t abs = prime0 + q, cs = seg_start + ((abs - seg_start) / chunk) * chunk;
                size_t ce = cs + chunk; if (ce > seg_end) ce = seg_end;
                res[q - q_data] = search_pos(cfg, lazy, abs, prime0, pre, cs, ce, cross, probe_k > 0 ? (unsigned)probe_k : 0, &walk[l], &rnd[l]);
                if (walk[l] > longest) longest = walk[l];
                data_positions++;
            }
            if (q0 + 32 <= q_data) continue;
            if (probe_k > 0) {
                /* the greedy parse over the probe results visits some positions of this batch: those are searched in full */
                const size_t batch_end = prime0 + q0 + 32 < seg_end ? prime0 + q0 + 32 : seg_end;
                for (unsigned l = 0; l < 32; l++) probe_cand += walk[l];
                while (pp1 < batch_end) {
                    const size_t cs = seg_start + ((pp1 - seg_start) / chunk) * chunk;
                    size_t ce = cs + chunk; if (ce > seg_end) ce = seg_end;
                    const unsigned L1 = (res[pp1 - seg_start] >> 15) & 0x1ffu;
                    unsigned w = 0, r8 = 0;
                    res[pp1 - seg_start] = search_pos(cfg, lazy, pp1, prime0, pre, cs, ce, cross, 0, &w, &r8);
                    full_positions++; full_cand += w;
                    cw[ncw++] = w;
                    if (ncw == 32) { unsigned m = 0; for (unsigned i = 0; i < 32; i++) if (cw[i] > m) m = cw[i]; full_simt += m; full_batches++; ncw = 0; }
                    pp1 += L1 >= 3 ? L1 : 1;           /* the probe's parse moves on by the PROBE's match */
                    if (pp1 > ce) pp1 = ce;
                }
                for (unsigned l = 0; l < 32; l++) walk[l] = rnd[l] = 0;   /* the probe walks are accounted separately */
                longest = 0;
            }
            if (stop_active && longest > (unsigned)stop_after) {
                /* the iteration at which at most stop_active lanes are still walking */
                unsigned cap = (unsigned)stop_after;
                for (;; cap++) {
                    int active = 0;
                    for (unsigned l = 0; l < 32; l++) active += walk[l] > cap;
                    if (active <= stop_active) break;
                }
                if (cap < longest) {
                    stopped++;
                    longest = 0;
                    for (unsigned l = 0; l < 32; l++) {
                        const size_t q = q0 + l;
                        if (q < q_data || q >= n) continue;
                        const size_t abs = prime0 + q, cs = seg_start + ((abs - seg_start) / chunk) * chunk;
                        size_t ce = cs + chunk; if (ce > seg_end) ce = seg_end;
                        if (walk[l] > cap) res[q - q_data] = search_pos(cfg, lazy, abs, prime0, pre, cs, ce, cross, cap, &walk[l], &rnd[l]);
                        if (walk[l] > longest) longest = walk[l];
                    }
                }
            }
            unsigned heavy = 0;
            for (unsigned l = 0; l < 32; l++) { cand_total += walk[l]; rounds_total += rnd[l]; if (walk[l] + rnd[l] > heavy) heavy = walk[l] + rnd[l]; }
            simt_total += longest;
            simt_heavy += heavy;
            if (heavy > step_heavy) step_heavy = heavy;
            if ((q0 + 32) % kStep == 0 || q0 + 32 >= n) { step_heavy_total += step_heavy; steps++; step_heavy = 0; }   /* the CTA's barrier */
            batches++;
            int b = 0; while ((1u << b) <= longest && b < 15) b++;
            hist[b]++;
        }
        PARSE_UNTIL(seg_end);
    }
    printf("bytes_in %zu  bytes_out %llu  ratio %.5f  blocks %llu  symbols %llu\n", total, (unsigned long long)((bits + 7) / 8),
           (double)((bits + 7) / 8) / (double)total, (unsigned long long)nblocks, (unsigned long long)nsym_total);
    printf("candidates %llu (%.2f per position)  batches %llu  simt_iterations %llu (%.2f per batch)  lanes_active %.2f of 32  batches_stopped %llu\n",
           (unsigned long long)cand_total, (double)cand_total / (double)total, (unsigned long long)batches, (unsigned long long)simt_total,
           (double)simt_total / (double)batches, simt_total ? (double)cand_total / (double)simt_total : 0.0, (unsigned long long)stopped);
    printf("compare_rounds %llu (%.2f per position)  simt_walk_plus_rounds %.2f per batch (the lane with most candidates + 8-byte compare rounds)\n",
           (unsigned long long)rounds_total, (double)rounds_total / (double)total, (double)simt_heavy / (double)batches);
    printf("steps %llu  heaviest batch of a 960-position step (what the CTA's barrier waits for): %.2f\n", (unsigned long long)steps,
           (double)step_heavy_total / (double)steps);
    if (probe_k > 0)
        printf("probe=%d: probe candidates %.2f per position; searched in full %.1f %% of the positions, %.2f candidates each, "
               "compacted batches: longest walk %.2f\n", probe_k, (double)probe_cand / (double)data_positions,
               100.0 * (double)full_positions / (double)data_positions, full_positions ? (double)full_cand / (double)full_positions : 0.0,
               full_batches ? (double)full_simt / (double)full_batches : 0.0);
    printf("longest walk per batch, buckets [0] [1] [2-3] [4-7] ...:");
    for (int i = 0; i < 16; i++) printf(" %llu", (unsigned long long)hist[i]);
    printf("\n");
    if (fd) fclose(fd);
    free(res); free(sym); free(data);
    return 0;
}
