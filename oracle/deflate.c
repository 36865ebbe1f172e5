/*
 * deflate.c -- oracle (test infrastructure): CPU restatement of the reference encoder for levels
 * 0..9 and every strategy, windowBits 15, memLevel 8 -- deflate/deflate.ts (CONFIGURATION_TABLE,
 * INSERT_STRING, fill_window/slide_hash, longest_match, deflate_stored, deflate_fast, deflate_slow,
 * deflate_huff, deflate_rle, deflateSetDictionary, header/trailer emission), deflate/trees.ts (build_tree, gen_bitlen,
 * gen_codes, scan_tree/send_tree, compress_block, _tr_flush_block, _tr_stored_block) and
 * deflate/utils.ts (_tr_tally_*, d_code).
 *
 * The reference streams through a 64 KiB sliding window; this restatement is one-shot: it keeps
 * the whole input addressable and works in absolute positions, carrying the window base `wbase`
 * that fill_window/slide_hash would have produced, so that hash chains, the NIL sentinel (window
 * index 0) and the MAX_DIST limit select exactly the same matches.  Used to obtain the
 * reference's compressed size per level (the ratio gate) and to cross-check block-format
 * decisions.  Not part of the product path.
 */
#include <stdlib.h>
#include <string.h>

#include "zs_oracle.h"

#define W_SIZE 32768u
#define W_MASK 32767u
#define HASH_MASK 32767u
#define MIN_MATCH 3
#define MAX_MATCH 258
#define MIN_LOOKAHEAD (MAX_MATCH + MIN_MATCH + 1)
#define MAX_DIST (W_SIZE - MIN_LOOKAHEAD)
#define TOO_FAR 4096
#define LIT_BUFSIZE 16384u            /* 1 << (memLevel + 6), deflate.ts:323 */
#define SYM_LIMIT (LIT_BUFSIZE - 1)   /* sym_end = (lit_bufsize - 1) * 3, deflate.ts:336 */
#define L_CODES 286
#define D_CODES 30
#define BL_CODES 19
#define HEAP_SIZE (2 * L_CODES + 1)
#define END_BLOCK 256

/* CONFIGURATION_TABLE, deflate.ts:86-103 */
typedef struct { int lazy_fn, good, lazy, nice, chain; } level_cfg;
static const level_cfg LEVELS[10] = {
    {0, 0, 0, 0, 0},      {0, 4, 4, 8, 4},      {0, 4, 5, 16, 8},      {0, 4, 6, 32, 32},
    {1, 4, 4, 16, 16},    {1, 8, 16, 32, 32},   {1, 8, 16, 128, 128},  {1, 8, 32, 128, 256},
    {1, 32, 128, 258, 1024}, {1, 32, 258, 258, 4096}};

/* ---- static tables (RFC 1951 3.2.5/3.2.6; deflate/constants.ts, trees-util.ts) -------------- */
static uint8_t LEN_CODE[256];      /* match length - 3 -> length code 0..28 */
static uint16_t LEN_BASE[29];
static uint8_t LEN_XBITS[29];
static uint16_t DIST_BASE[30];
static uint8_t DIST_XBITS[30];
static uint8_t DCODE_LO[256], DCODE_HI[256]; /* dist-1 < 256 ; (dist-1) >> 7 */
static uint8_t BL_XBITS[19] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 3, 7};
static const uint8_t BL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
static uint16_t ST_LLEN[288], ST_LCODE[288], ST_DLEN[30], ST_DCODE[30];
static int tables_ready = 0;

static unsigned bit_reverse(unsigned code, int len) {
    unsigned r = 0;
    while (len-- > 0) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

/* gen_codes, trees.ts:54-76: canonical codes, emitted LSB first hence bit-reversed */
static void gen_codes(const uint16_t* len, uint16_t* code, int max_code, const uint16_t* bl_count) {
    uint16_t next[16];
    unsigned c = 0;
    for (int b = 1; b <= 15; b++) { c = (c + bl_count[b - 1]) << 1; next[b] = (uint16_t)c; }
    for (int n = 0; n <= max_code; n++) {
        if (len[n] == 0) continue;
        code[n] = (uint16_t)bit_reverse(next[len[n]]++, len[n]);
    }
}

__attribute__((constructor)) static void tables_build(void) {
    unsigned base = 0;
    for (int c = 0; c < 28; c++) {
        int xb = c < 8 ? 0 : (c - 4) / 4;
        LEN_BASE[c] = (uint16_t)base;
        LEN_XBITS[c] = (uint8_t)xb;
        for (unsigned k = 0; k < (1u << xb); k++) LEN_CODE[base + k] = (uint8_t)c;
        base += 1u << xb;
    }
    LEN_BASE[28] = 0; LEN_XBITS[28] = 0; LEN_CODE[255] = 28; /* length 258 has its own code */
    base = 0;
    for (int c = 0; c < 30; c++) {
        int xb = c < 4 ? 0 : (c - 2) / 2;
        DIST_BASE[c] = (uint16_t)base;
        DIST_XBITS[c] = (uint8_t)xb;
        for (unsigned k = 0; k < (1u << xb); k++) {
            unsigned d = base + k;
            if (d < 256) DCODE_LO[d] = (uint8_t)c;
            if ((d & 127u) == 0 && d >= 256) DCODE_HI[d >> 7] = (uint8_t)c;
        }
        base += 1u << xb;
    }
    uint16_t cnt[16] = {0};
    for (int n = 0; n < 288; n++) {
        ST_LLEN[n] = (uint16_t)(n < 144 ? 8 : n < 256 ? 9 : n < 280 ? 7 : 8);
        cnt[ST_LLEN[n]]++;
    }
    gen_codes(ST_LLEN, ST_LCODE, 287, cnt);
    for (int n = 0; n < 30; n++) { ST_DLEN[n] = 5; ST_DCODE[n] = (uint16_t)bit_reverse((unsigned)n, 5); }
    tables_ready = 1;
}

/* d_code, deflate/utils.ts:87 */
static inline unsigned d_code(unsigned dist) { return dist < 256 ? DCODE_LO[dist] : DCODE_HI[dist >> 7]; }

/* ---- Huffman construction ------------------------------------------------------------------ */
typedef struct {
    uint16_t freq[HEAP_SIZE], len[HEAP_SIZE], code[HEAP_SIZE], dad[HEAP_SIZE];
    int max_code;
} tree_t;

typedef struct {
    const uint16_t* static_len; /* NULL for the bit-length tree */
    const uint8_t* xbits;
    int xbase, elems, max_length;
} tree_kind;

static const tree_kind KIND_L = {ST_LLEN, LEN_XBITS, 257, L_CODES, 15};
static const tree_kind KIND_D = {ST_DLEN, DIST_XBITS, 0, D_CODES, 15};
static const tree_kind KIND_BL = {NULL, BL_XBITS, 0, BL_CODES, 7};

typedef struct {
    int heap[HEAP_SIZE + 1];
    uint8_t depth[HEAP_SIZE];
    int heap_len, heap_max;
    uint16_t bl_count[16];
    uint32_t opt_len, static_len;
} build_ctx;

/* smaller, trees.ts:163-165 */
static inline int node_smaller(const tree_t* t, const build_ctx* b, int n, int m) {
    return t->freq[n] < t->freq[m] || (t->freq[n] == t->freq[m] && b->depth[n] <= b->depth[m]);
}

/* pqdownheap, trees.ts:167-185 */
static void sift_down(const tree_t* t, build_ctx* b, int k) {
    int v = b->heap[k];
    int j = k << 1;
    while (j <= b->heap_len) {
        if (j < b->heap_len && node_smaller(t, b, b->heap[j + 1], b->heap[j])) j++;
        if (node_smaller(t, b, v, b->heap[j])) break;
        b->heap[k] = b->heap[j];
        k = j;
        j <<= 1;
    }
    b->heap[k] = v;
}

/* gen_bitlen, trees.ts:187-259 */
static void gen_bitlen(tree_t* t, build_ctx* b, const tree_kind* kd) {
    int h, n, m, bits, overflow = 0;
    for (bits = 0; bits <= 15; bits++) b->bl_count[bits] = 0;
    t->len[b->heap[b->heap_max]] = 0;
    for (h = b->heap_max + 1; h < HEAP_SIZE; h++) {
        n = b->heap[h];
        bits = t->len[t->dad[n]] + 1;
        if (bits > kd->max_length) { bits = kd->max_length; overflow++; }
        t->len[n] = (uint16_t)bits;
        if (n > t->max_code) continue; /* not a leaf */
        b->bl_count[bits]++;
        int xb = n >= kd->xbase ? kd->xbits[n - kd->xbase] : 0;
        uint32_t f = t->freq[n];
        b->opt_len += f * (uint32_t)(bits + xb);
        if (kd->static_len) b->static_len += f * (uint32_t)(kd->static_len[n] + xb);
    }
    if (overflow == 0) return;
    do {
        bits = kd->max_length - 1;
        while (b->bl_count[bits] == 0) bits--;
        b->bl_count[bits]--;
        b->bl_count[bits + 1] += 2;
        b->bl_count[kd->max_length]--;
        overflow -= 2;
    } while (overflow > 0);
    for (bits = kd->max_length; bits != 0; bits--) {
        n = b->bl_count[bits];
        while (n != 0) {
            m = b->heap[--h];
            if (m > t->max_code) continue;
            if (t->len[m] != (unsigned)bits) {
                b->opt_len += (uint32_t)(((long)bits - (long)t->len[m]) * (long)t->freq[m]);
                t->len[m] = (uint16_t)bits;
            }
            n--;
        }
    }
}

/* build_tree, trees.ts:261-316 */
static void build_tree(tree_t* t, build_ctx* b, const tree_kind* kd) {
    int n, m, node, max_code = -1;
    b->heap_len = 0;
    b->heap_max = HEAP_SIZE;
    for (n = 0; n < kd->elems; n++) {
        if (t->freq[n] != 0) { b->heap[++b->heap_len] = max_code = n; b->depth[n] = 0; }
        else t->len[n] = 0;
    }
    while (b->heap_len < 2) { /* force at least two codes of non-zero frequency */
        node = b->heap[++b->heap_len] = (max_code < 2 ? ++max_code : 0);
        t->freq[node] = 1;
        b->depth[node] = 0;
        b->opt_len--;
        if (kd->static_len) b->static_len -= kd->static_len[node];
    }
    t->max_code = max_code;
    for (n = b->heap_len / 2; n >= 1; n--) sift_down(t, b, n);
    node = kd->elems;
    do {
        n = b->heap[1];                       /* pqremove */
        b->heap[1] = b->heap[b->heap_len--];
        sift_down(t, b, 1);
        m = b->heap[1];
        b->heap[--b->heap_max] = n;
        b->heap[--b->heap_max] = m;
        t->freq[node] = (uint16_t)(t->freq[n] + t->freq[m]);
        b->depth[node] = (uint8_t)((b->depth[n] >= b->depth[m] ? b->depth[n] : b->depth[m]) + 1);
        t->dad[n] = t->dad[m] = (uint16_t)node;
        b->heap[1] = node++;
        sift_down(t, b, 1);
    } while (b->heap_len >= 2);
    b->heap[--b->heap_max] = b->heap[1];
    gen_bitlen(t, b, kd);
    gen_codes(t->len, t->code, max_code, b->bl_count);
}
/* note: the reference's gen_codes reads bl_count_arr[bits - 1] with the array passed unshifted
 * (trees.ts:64), i.e. next_code[bits] = (code + bl_count[bits-1]) << 1 -- the RFC rule. */

int zo_build_tree(int kind, uint16_t* freq, uint16_t* len_out, uint16_t* code_out, uint32_t* opt_len,
                  uint32_t* static_len) {
    if (!tables_ready) tables_build();
    const tree_kind* kd = kind == 0 ? &KIND_L : kind == 1 ? &KIND_D : &KIND_BL;
    tree_t* t = (tree_t*)calloc(1, sizeof(tree_t));
    build_ctx* b = (build_ctx*)calloc(1, sizeof(build_ctx));
    memcpy(t->freq, freq, sizeof(uint16_t) * (size_t)kd->elems);
    build_tree(t, b, kd);
    memcpy(freq, t->freq, sizeof(uint16_t) * (size_t)kd->elems);
    memcpy(len_out, t->len, sizeof(uint16_t) * (size_t)kd->elems);
    memcpy(code_out, t->code, sizeof(uint16_t) * (size_t)kd->elems);
    if (opt_len) *opt_len = b->opt_len;
    if (static_len) *static_len = b->static_len;
    int mc = t->max_code;
    free(t);
    free(b);
    return mc;
}

/* ---- encoder state --------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* win;   /* virtual buffer: [dictionary][input] */
    size_t total;         /* bytes in the virtual buffer */
    size_t strstart, fill, wbase, block_start;
    size_t insert;
    uint32_t head[W_SIZE], prev[W_SIZE];
    int level, strategy;
    size_t match_start;
    unsigned match_length, prev_length;
    size_t prev_match;
    int match_available;
    /* symbol buffer, deflate/utils.ts:55-81 */
    uint16_t sym_dist[LIT_BUFSIZE];
    uint8_t sym_lc[LIT_BUFSIZE];
    unsigned sym_next;
    tree_t lt, dt, blt;
    build_ctx bc;
    /* bit writer */
    uint8_t* out;
    size_t out_cap, out_pos;
    uint64_t acc;
    unsigned acc_bits;
    int overflow;
} enc_t;

/* send_bits, trees.ts:78-88 (the 16-bit staging buffer is an implementation detail; the emitted
 * byte sequence is the LSB-first concatenation) */
static void put_bits(enc_t* e, unsigned value, unsigned nbits) {
    e->acc |= (uint64_t)value << e->acc_bits;
    e->acc_bits += nbits;
    while (e->acc_bits >= 8) {
        if (e->out_pos < e->out_cap) e->out[e->out_pos] = (uint8_t)e->acc; else e->overflow = 1;
        e->out_pos++;
        e->acc >>= 8;
        e->acc_bits -= 8;
    }
}
/* bi_windup, trees.ts:43-52 */
static void byte_align(enc_t* e) { if (e->acc_bits) put_bits(e, 0, 8 - e->acc_bits); }
static void put_byte(enc_t* e, unsigned b) { put_bits(e, b & 0xffu, 8); }

/* init_block, trees.ts:90-103 */
static void init_block(enc_t* e) {
    memset(e->lt.freq, 0, sizeof(e->lt.freq));
    memset(e->dt.freq, 0, sizeof(e->dt.freq));
    memset(e->blt.freq, 0, sizeof(e->blt.freq));
    e->lt.freq[END_BLOCK] = 1;
    e->bc.opt_len = e->bc.static_len = 0;
    e->sym_next = 0;
}

/* _tr_tally_lit / _tr_tally_dist, deflate/utils.ts:55-81 */
static int tally_lit(enc_t* e, unsigned c) {
    e->sym_dist[e->sym_next] = 0;
    e->sym_lc[e->sym_next++] = (uint8_t)c;
    e->lt.freq[c]++;
    return e->sym_next == SYM_LIMIT;
}
static int tally_dist(enc_t* e, unsigned dist, unsigned lc) {
    e->sym_dist[e->sym_next] = (uint16_t)dist;
    e->sym_lc[e->sym_next++] = (uint8_t)lc;
    dist--;
    e->lt.freq[LEN_CODE[lc] + 257]++;
    e->dt.freq[d_code(dist)]++;
    return e->sym_next == SYM_LIMIT;
}

/* scan_tree, trees.ts:318-363 */
static void scan_tree(enc_t* e, tree_t* t, int max_code) {
    int prevlen = -1, curlen, nextlen = t->len[0], count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    t->len[max_code + 1] = 0xffff;
    for (int n = 0; n <= max_code; n++) {
        curlen = nextlen;
        nextlen = t->len[n + 1];
        if (++count < max_count && curlen == nextlen) continue;
        else if (count < min_count) e->blt.freq[curlen] += (uint16_t)count;
        else if (curlen != 0) { if (curlen != prevlen) e->blt.freq[curlen]++; e->blt.freq[16]++; }
        else if (count <= 10) e->blt.freq[17]++;
        else e->blt.freq[18]++;
        count = 0;
        prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

/* send_tree, trees.ts:365-414 */
static void send_tree(enc_t* e, tree_t* t, int max_code) {
    int prevlen = -1, curlen, nextlen = t->len[0], count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    for (int n = 0; n <= max_code; n++) {
        curlen = nextlen;
        nextlen = t->len[n + 1];
        if (++count < max_count && curlen == nextlen) continue;
        else if (count < min_count) {
            do put_bits(e, e->blt.code[curlen], e->blt.len[curlen]); while (--count != 0);
        } else if (curlen != 0) {
            if (curlen != prevlen) { put_bits(e, e->blt.code[curlen], e->blt.len[curlen]); count--; }
            put_bits(e, e->blt.code[16], e->blt.len[16]);
            put_bits(e, (unsigned)(count - 3), 2);
        } else if (count <= 10) {
            put_bits(e, e->blt.code[17], e->blt.len[17]);
            put_bits(e, (unsigned)(count - 3), 3);
        } else {
            put_bits(e, e->blt.code[18], e->blt.len[18]);
            put_bits(e, (unsigned)(count - 11), 7);
        }
        count = 0;
        prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

/* compress_block, trees.ts:476-520 */
static void compress_block(enc_t* e, const uint16_t* lcode, const uint16_t* llen, const uint16_t* dcode,
                           const uint16_t* dlen) {
    for (unsigned i = 0; i < e->sym_next; i++) {
        unsigned dist = e->sym_dist[i], lc = e->sym_lc[i];
        if (dist == 0) {
            put_bits(e, lcode[lc], llen[lc]);
        } else {
            unsigned code = LEN_CODE[lc];
            put_bits(e, lcode[code + 257], llen[code + 257]);
            if (LEN_XBITS[code]) put_bits(e, lc - LEN_BASE[code], LEN_XBITS[code]);
            dist--;
            code = d_code(dist);
            put_bits(e, dcode[code], dlen[code]);
            if (DIST_XBITS[code]) put_bits(e, dist - DIST_BASE[code], DIST_XBITS[code]);
        }
    }
    put_bits(e, lcode[END_BLOCK], llen[END_BLOCK]);
}

/* _tr_stored_block, trees.ts:449-464 */
static void stored_block(enc_t* e, const uint8_t* buf, size_t len, int last) {
    put_bits(e, (unsigned)last, 3);
    byte_align(e);
    put_byte(e, (unsigned)len); put_byte(e, (unsigned)(len >> 8));
    put_byte(e, (unsigned)~len); put_byte(e, (unsigned)(~len >> 8));
    for (size_t i = 0; i < len; i++) put_byte(e, buf[i]);
}

/* _tr_flush_block, trees.ts:544-590 (+ build_bl_tree :416-432, send_all_trees :434-447) */
static void flush_block(enc_t* e, int last) {
    size_t stored_len = e->strstart - e->block_start;
    int can_store = e->block_start >= e->wbase; /* C zlib passes NULL once the block start slid out */
    build_tree(&e->lt, &e->bc, &KIND_L);
    build_tree(&e->dt, &e->bc, &KIND_D);
    scan_tree(e, &e->lt, e->lt.max_code);
    scan_tree(e, &e->dt, e->dt.max_code);
    build_tree(&e->blt, &e->bc, &KIND_BL);
    int max_blindex;
    for (max_blindex = BL_CODES - 1; max_blindex >= 3; max_blindex--)
        if (e->blt.len[BL_ORDER[max_blindex]] != 0) break;
    e->bc.opt_len += 3u * ((unsigned)max_blindex + 1) + 5 + 5 + 4;
    size_t opt_lenb = (e->bc.opt_len + 3 + 7) >> 3;
    size_t static_lenb = (e->bc.static_len + 3 + 7) >> 3;
    if (static_lenb <= opt_lenb || e->strategy == ZO_FIXED) opt_lenb = static_lenb; /* trees.ts:567 */

    if (stored_len + 4 <= opt_lenb && can_store) {
        stored_block(e, e->win + e->block_start, stored_len, last);
    } else if (static_lenb == opt_lenb) {
        put_bits(e, (1u << 1) + (unsigned)last, 3);
        compress_block(e, ST_LCODE, ST_LLEN, ST_DCODE, ST_DLEN);
    } else {
        put_bits(e, (2u << 1) + (unsigned)last, 3);
        put_bits(e, (unsigned)(e->lt.max_code + 1 - 257), 5);
        put_bits(e, (unsigned)(e->dt.max_code + 1 - 1), 5);
        put_bits(e, (unsigned)(max_blindex + 1 - 4), 4);
        for (int r = 0; r <= max_blindex; r++) put_bits(e, e->blt.len[BL_ORDER[r]], 3);
        send_tree(e, &e->lt, e->lt.max_code);
        send_tree(e, &e->dt, e->dt.max_code);
        compress_block(e, e->lt.code, e->lt.len, e->dt.code, e->dt.len);
    }
    init_block(e);
    if (last) byte_align(e);
    e->block_start = e->strstart;
}

/* ---- LZ77 ------------------------------------------------------------------------------------ */
static inline unsigned hash3(const uint8_t* p) { /* three UPDATE_HASH steps, deflate.ts:109-111 */
    return (((unsigned)p[0] << 10) ^ ((unsigned)p[1] << 5) ^ p[2]) & HASH_MASK;
}

/* INSERT_STRING, deflate.ts:113-118 */
static inline size_t insert_string(enc_t* e, size_t pos) {
    unsigned h = hash3(e->win + pos);
    size_t hd = e->head[h];
    e->prev[pos & W_MASK] = (uint32_t)hd;
    e->head[h] = (uint32_t)pos;
    return hd;
}

/* fill_window, deflate.ts:166-236 -- only its bookkeeping: how far the window has been filled,
 * when it slides (slide_hash :125-141 becomes "entries <= wbase are NIL"), and the deferred
 * `insert` positions. */
static void fill_window(enc_t* e) {
    do {
        size_t more = 2 * W_SIZE - (e->fill - e->wbase);
        if (e->strstart - e->wbase >= W_SIZE + MAX_DIST) {
            e->wbase += W_SIZE;
            more += W_SIZE;
            if (e->insert > e->strstart - e->wbase) e->insert = e->strstart - e->wbase;
        }
        size_t avail = e->total - e->fill;
        if (avail == 0) break;
        size_t n = avail < more ? avail : more;
        e->fill += n;
        size_t lookahead = e->fill - e->strstart;
        if (lookahead + e->insert >= MIN_MATCH) {
            size_t str = e->strstart - e->insert;
            while (e->insert) {
                insert_string(e, str);
                str++;
                e->insert--;
                if (lookahead + e->insert < MIN_MATCH) break;
            }
        }
    } while (e->fill - e->strstart < MIN_LOOKAHEAD && e->total != e->fill);
}

/* longest_match, deflate.ts:1053-1115 */
static unsigned longest_match(enc_t* e, size_t cur_match) {
    const level_cfg* cfg = &LEVELS[e->level];
    unsigned chain = (unsigned)cfg->chain;
    const uint8_t* win = e->win;
    const size_t scan = e->strstart;
    unsigned best_len = e->prev_length;
    size_t lookahead = e->fill - e->strstart;
    unsigned nice = (unsigned)cfg->nice;
    size_t rel = e->strstart - e->wbase;
    size_t limit = rel > MAX_DIST ? e->strstart - MAX_DIST : e->wbase;
    unsigned max_cmp = MAX_MATCH < lookahead ? MAX_MATCH : (unsigned)lookahead;
    uint8_t scan_end1 = win[scan + best_len - 1], scan_end = win[scan + best_len];
    if (best_len >= (unsigned)cfg->good) chain >>= 2;
    if (nice > lookahead) nice = (unsigned)lookahead;
    do {
        size_t m = cur_match;
        if (win[m + best_len] != scan_end || win[m + best_len - 1] != scan_end1 || win[m] != win[scan] ||
            win[m + 1] != win[scan + 1])
            continue;
        unsigned k = 2;
        while (k < max_cmp && win[scan + k] == win[m + k]) k++;
        if (k > best_len) {
            e->match_start = cur_match;
            best_len = k;
            if (k >= nice) break;
            scan_end1 = win[scan + best_len - 1];
            scan_end = win[scan + best_len];
        }
    } while ((cur_match = e->prev[cur_match & W_MASK]) > limit && --chain != 0);
    return best_len <= lookahead ? best_len : (unsigned)lookahead;
}

/* deflate_fast, deflate.ts:1281-1350, driven to the end of the input (flush != Z_NO_FLUSH) */
static void run_fast(enc_t* e) {
    const level_cfg* cfg = &LEVELS[e->level];
    for (;;) {
        if (e->fill - e->strstart < MIN_LOOKAHEAD) {
            fill_window(e);
            if (e->fill == e->strstart) break;
        }
        size_t lookahead = e->fill - e->strstart;
        size_t hash_head = e->wbase; /* NIL */
        if (lookahead >= MIN_MATCH) hash_head = insert_string(e, e->strstart);
        if (hash_head > e->wbase && e->strstart - hash_head <= MAX_DIST) {
            e->prev_length = MIN_MATCH - 1; /* deflate_fast never raises prev_length */
            e->match_length = longest_match(e, hash_head);
        }
        int bflush;
        if (e->match_length >= MIN_MATCH) {
            bflush = tally_dist(e, (unsigned)(e->strstart - e->match_start), e->match_length - MIN_MATCH);
            lookahead -= e->match_length;
            if (e->match_length <= (unsigned)cfg->lazy && lookahead >= MIN_MATCH) {
                e->match_length--;
                do { e->strstart++; insert_string(e, e->strstart); } while (--e->match_length != 0);
                e->strstart++;
            } else {
                e->strstart += e->match_length;
                e->match_length = 0;
            }
        } else {
            bflush = tally_lit(e, e->win[e->strstart]);
            e->strstart++;
        }
        if (bflush) flush_block(e, 0);
    }
}

/* deflate_slow, deflate.ts:1352-1448, driven to the end of the input */
static void run_slow(enc_t* e) {
    const level_cfg* cfg = &LEVELS[e->level];
    for (;;) {
        if (e->fill - e->strstart < MIN_LOOKAHEAD) {
            fill_window(e);
            if (e->fill == e->strstart) break;
        }
        size_t lookahead = e->fill - e->strstart;
        size_t hash_head = e->wbase;
        if (lookahead >= MIN_MATCH) hash_head = insert_string(e, e->strstart);
        e->prev_length = e->match_length;
        e->prev_match = e->match_start;
        e->match_length = MIN_MATCH - 1;
        if (hash_head > e->wbase && e->prev_length < (unsigned)cfg->lazy &&
            e->strstart - hash_head <= MAX_DIST) {
            e->match_length = longest_match(e, hash_head);
            if (e->match_length <= 5 && (e->strategy == ZO_FILTERED || /* deflate.ts:1386-1392 */
                                         (e->match_length == MIN_MATCH && e->strstart - e->match_start > TOO_FAR)))
                e->match_length = MIN_MATCH - 1;
        }
        if (e->prev_length >= MIN_MATCH && e->match_length <= e->prev_length) {
            size_t max_insert = e->strstart + lookahead - MIN_MATCH;
            int bflush = tally_dist(e, (unsigned)(e->strstart - 1 - e->prev_match), e->prev_length - MIN_MATCH);
            e->prev_length -= 2;
            do {
                if (++e->strstart <= max_insert) insert_string(e, e->strstart);
            } while (--e->prev_length != 0);
            e->match_available = 0;
            e->match_length = MIN_MATCH - 1;
            e->strstart++;
            if (bflush) flush_block(e, 0);
        } else if (e->match_available) {
            if (tally_lit(e, e->win[e->strstart - 1])) flush_block(e, 0);
            e->strstart++;
        } else {
            e->match_available = 1;
            e->strstart++;
        }
    }
    if (e->match_available) {
        tally_lit(e, e->win[e->strstart - 1]);
        e->match_available = 0;
    }
}

/* deflate_huff, deflate.ts:1525-1560, driven to the end of the input: literals only */
static void run_huff(enc_t* e) {
    for (;;) {
        if (e->fill == e->strstart) {
            fill_window(e);
            if (e->fill == e->strstart) break;
        }
        e->match_length = 0;
        int bflush = tally_lit(e, e->win[e->strstart]);
        e->strstart++;
        if (bflush) flush_block(e, 0);
    }
}

/* deflate_rle as C zlib 1.3 defines it (deflate.c deflate_rle): matches at distance 1 only.  The
 * reference's port compares the previous byte with the scan INDEX instead of the byte at that
 * index (deflate.ts:1467-1469, `prev == ++scan`), which can never hold three times in a row, so
 * the reference itself emits literals only for Z_RLE -- byte for byte what run_huff produces
 * (zo_deflate_oneshot2's `rle_like_reference`).  The engine implements the intended algorithm,
 * so this is the one it is compared with. */
static void run_rle(enc_t* e) {
    for (;;) {
        if (e->fill - e->strstart <= MAX_MATCH) {
            fill_window(e);
            if (e->fill == e->strstart) break;
        }
        size_t lookahead = e->fill - e->strstart;
        unsigned ml = 0;
        if (lookahead >= MIN_MATCH && e->strstart > 0) {
            const uint8_t* w = e->win;
            const uint8_t prev = w[e->strstart - 1];
            while (ml < MAX_MATCH && w[e->strstart + ml] == prev) ml++;   /* the 8-way unrolled scan */
            if (ml < MIN_MATCH) ml = 0;
            if (ml > lookahead) ml = (unsigned)lookahead;
        }
        int bflush;
        if (ml >= MIN_MATCH) {
            bflush = tally_dist(e, 1, ml - MIN_MATCH);
            e->strstart += ml;
        } else {
            bflush = tally_lit(e, e->win[e->strstart]);
            e->strstart++;
        }
        if (bflush) flush_block(e, 0);
    }
}

/* deflate_stored, deflate.ts:1140-1279, for one call that holds the whole input and an output
 * buffer of at least deflateBound bytes: the first loop (:1146-1197) then copies straight from the
 * input in blocks of MAX_STORED = 65535 bytes -- `have` never limits a block, and a shorter block is
 * only ever the last piece (`len == left + avail_in`) -- and marks the block that takes the last
 * byte as final under Z_FINISH; an empty input still gets its empty final block (:1160-1163). */
static void run_stored(enc_t* e, int finish) {
    size_t pos = e->strstart;
    for (;;) {
        size_t rest = e->total - pos;
        size_t len = rest > 65535 ? 65535 : rest;
        if (len == 0 && !finish) break;
        int last = finish && len == rest;
        stored_block(e, e->win + pos, len, last);
        pos += len;
        if (last || pos == e->total) break;
    }
    e->strstart = e->block_start = pos;
}

/* deflateBound, deflate.ts:615-674 for the default windowBits 15 / memLevel 8 state */
size_t zo_deflate_bound(size_t n, int wrap) {
    size_t wraplen = wrap == 0 ? 0 : wrap == 1 ? 6 : 18;
    return n + (n >> 12) + (n >> 14) + (n >> 25) + 13 - 6 + wraplen;
}

int64_t zo_deflate_oneshot(const uint8_t* in, size_t in_len, int level, int wrap, const uint8_t* dict,
                           size_t dict_len, int flush, uint8_t* out, size_t out_cap) {
    if (level == 0) return ZO_STREAM_ERROR; /* the historical entry point: levels 1..9 */
    return zo_deflate_oneshot2(in, in_len, level, ZO_DEFAULT_STRATEGY, 0, wrap, dict, dict_len, flush, out, out_cap);
}

int64_t zo_deflate_oneshot2(const uint8_t* in, size_t in_len, int level, int strategy, int rle_like_reference,
                            int wrap, const uint8_t* dict, size_t dict_len, int flush, uint8_t* out,
                            size_t out_cap) {
    if (level == -1) level = 6;
    if (level < 0 || level > 9 || wrap < 0 || wrap > 2 || strategy < 0 || strategy > ZO_FIXED) return ZO_STREAM_ERROR;
    if (flush != ZO_FINISH && flush != ZO_SYNC_FLUSH) return ZO_STREAM_ERROR;
    if (dict_len && wrap == 2) return ZO_STREAM_ERROR; /* deflateSetDictionary, deflate.ts:373 */
    if (!tables_ready) tables_build();
    enc_t* e = (enc_t*)calloc(1, sizeof(enc_t));
    if (!e) return ZO_MEM_ERROR;
    /* DICTID is the adler32 of the whole dictionary (deflate.ts:377-379); only its last w_size
     * bytes are loaded (deflate.ts:383-392) */
    uint32_t dict_id = dict_len ? zo_adler32(1u, dict, dict_len) : 0;
    if (dict_len > W_SIZE) { dict += dict_len - W_SIZE; dict_len = W_SIZE; }
    uint8_t* vbuf = (uint8_t*)malloc(dict_len + in_len + MAX_MATCH + 8);
    if (!vbuf) { free(e); return ZO_MEM_ERROR; }
    if (dict_len) memcpy(vbuf, dict, dict_len);
    if (in_len) memcpy(vbuf + dict_len, in, in_len);
    memset(vbuf + dict_len + in_len, 0, MAX_MATCH + 8);
    e->win = vbuf;
    e->level = level;
    e->strategy = strategy;
    e->out = out;
    e->out_cap = out_cap;
    e->match_length = e->prev_length = MIN_MATCH - 1;
    init_block(e);

    /* headers, deflate.ts:750-832 */
    if (wrap == 1) {
        unsigned header = (8u + (7u << 4)) << 8;
        unsigned lf = (strategy >= ZO_HUFFMAN_ONLY || level < 2) ? 0 : level < 6 ? 1 : level == 6 ? 2 : 3; /* deflate.ts:757-766 */
        header |= lf << 6;
        if (dict_len) header |= 0x20;
        header += 31 - header % 31;
        put_byte(e, header >> 8); put_byte(e, header);
        if (dict_len) {
            uint32_t id = dict_id;
            put_byte(e, id >> 24); put_byte(e, id >> 16); put_byte(e, id >> 8); put_byte(e, id);
        }
    } else if (wrap == 2) {
        put_byte(e, 0x1f); put_byte(e, 0x8b); put_byte(e, 8);
        for (int i = 0; i < 5; i++) put_byte(e, 0);
        put_byte(e, level == 9 ? 2 : (strategy >= ZO_HUFFMAN_ONLY || level < 2) ? 4 : 0); /* deflate.ts:797 */
        put_byte(e, 255); /* OS_CODE, deflate/constants.ts:30 */
    }

    /* deflateSetDictionary, deflate.ts:367-424: load the dictionary, hash all but its last two
     * positions, leave those as deferred inserts */
    if (dict_len) {
        e->total = dict_len;
        fill_window(e);
        while (e->fill - e->strstart >= MIN_MATCH) {
            size_t n = e->fill - e->strstart - (MIN_MATCH - 1);
            size_t str = e->strstart;
            do { insert_string(e, str); str++; } while (--n);
            e->strstart = str;
            fill_window(e);
        }
        e->insert = e->fill - e->strstart;
        e->strstart = e->fill;
        e->block_start = e->strstart;
    }
    e->total = dict_len + in_len;

    /* deflate(), deflate.ts:917-926: which block function runs */
    const int stored = level == 0;
    if (stored) run_stored(e, flush == ZO_FINISH);
    else if (strategy == ZO_HUFFMAN_ONLY || (strategy == ZO_RLE && rle_like_reference)) run_huff(e);
    else if (strategy == ZO_RLE) run_rle(e);
    else if (LEVELS[level].lazy_fn) run_slow(e);
    else run_fast(e);

    if (stored) {
        if (flush != ZO_FINISH) stored_block(e, NULL, 0, 0);
    } else if (flush == ZO_FINISH) {
        flush_block(e, 1);
    } else {
        if (e->sym_next) flush_block(e, 0);
        stored_block(e, NULL, 0, 0); /* Z_SYNC_FLUSH marker, deflate.ts:945-946 */
    }
    /* trailers, deflate.ts:964-988 */
    if (flush == ZO_FINISH) {
        if (wrap == 1) {
            uint32_t a = zo_adler32(1u, in, in_len);
            put_byte(e, a >> 24); put_byte(e, a >> 16); put_byte(e, a >> 8); put_byte(e, a);
        } else if (wrap == 2) {
            uint32_t c = zo_crc32(0u, in, in_len);
            put_byte(e, c); put_byte(e, c >> 8); put_byte(e, c >> 16); put_byte(e, c >> 24);
            put_byte(e, (unsigned)in_len); put_byte(e, (unsigned)(in_len >> 8));
            put_byte(e, (unsigned)(in_len >> 16)); put_byte(e, (unsigned)(in_len >> 24));
        }
    }
    int64_t produced = e->overflow ? ZO_BUF_ERROR : (int64_t)e->out_pos;
    free(vbuf);
    free(e);
    return produced;
}
