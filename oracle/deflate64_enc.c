/* deflate64_enc.c -- TEST INFRASTRUCTURE ONLY (see zs_oracle.h): a small raw deflate64 encoder.
 *
 * The reference ships no deflate64 encoder (src/mod/streams.ts:220-233 offers "deflate64-raw" for
 * decompression only; its test data in test/data was produced elsewhere), so the parity tests that
 * need deflate64 streams beyond the ten fixtures -- 64 KiB window with distances above 32768, and
 * the length code 285 that carries 16 extra bits -- make them here.  The format is what the
 * reference's decoder accepts: RFC 1951 with the deflate64 changes of
 * src/mod/inflate/constants.ts:39-45 (code 285 = base 3 + 16 extra bits; distance codes 30 / 31 =
 * bases 32769 / 49153 + 14 extra bits).  One fixed-Huffman block per call (BTYPE 1), greedy
 * hash-chain matching; `max_len` caps match lengths (258 keeps to codes <= 284, 65538 allows 285).
 */
#include <stdlib.h>
#include <string.h>

#include "zs_oracle.h"

typedef struct { uint8_t* out; size_t cap, pos; uint64_t acc; unsigned n; int overflow; } bitw;

static void put(bitw* w, unsigned v, unsigned n) {
    w->acc |= (uint64_t)v << w->n;
    w->n += n;
    while (w->n >= 8) {
        if (w->pos < w->cap) w->out[w->pos++] = (uint8_t)w->acc; else w->overflow = 1;
        w->acc >>= 8;
        w->n -= 8;
    }
}
static unsigned rev(unsigned code, unsigned len) {
    unsigned r = 0;
    while (len--) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}
/* fixed literal/length code, RFC 1951 3.2.6 */
static void put_litlen(bitw* w, unsigned sym) {
    if (sym < 144) put(w, rev(0x30 + sym, 8), 8);
    else if (sym < 256) put(w, rev(0x190 + (sym - 144), 9), 9);
    else if (sym < 280) put(w, rev(sym - 256, 7), 7);
    else put(w, rev(0xC0 + (sym - 280), 8), 8);
}

static void put_length(bitw* w, unsigned len, int allow285) {
    /* codes 257..284 as in deflate: extra bits 0 x8, then 1..5 x4 each */
    if (len <= 257 || (len == 258 && !allow285)) {
        unsigned base = 3, i;
        if (len == 258) { /* deflate64 has no 0-extra-bit code for 258: 284 with extra 31 */
            put_litlen(w, 284); put(w, 31, 5); return;
        }
        for (i = 0; i < 28; i++) {
            unsigned eb = i < 8 ? 0 : (i - 4) / 4;
            if (len < base + (1u << eb)) { put_litlen(w, 257 + i); put(w, len - base, eb); return; }
            base += 1u << eb;
        }
    }
    put_litlen(w, 285);              /* base 3, 16 extra bits */
    put(w, len - 3, 16);
}
static void put_distance(bitw* w, unsigned dist) {
    unsigned base = 1, i;
    for (i = 0; i < 32; i++) {
        unsigned eb = i < 4 ? 0 : (i - 2) / 2;   /* codes 30, 31: 14 extra bits */
        if (dist < base + (1u << eb)) { put(w, rev(i, 5), 5); put(w, dist - base, eb); return; }
        base += 1u << eb;
    }
}

long zo_deflate64_encode(const uint8_t* in, size_t n, unsigned max_len, int final, uint8_t* out, size_t cap) {
    enum { HB = 16, WIN = 65536, CHAIN = 64 };
    bitw w = {out, cap, 0, 0, 0, 0};
    int32_t* head = (int32_t*)malloc(sizeof(int32_t) << HB);
    int32_t* prev = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
    if (!head || !prev) { free(head); free(prev); return -1; }
    memset(head, 0xff, sizeof(int32_t) << HB);
    if (max_len < 3) max_len = 3;
    if (max_len > 65538) max_len = 65538;
    put(&w, final ? 1 : 0, 1);
    put(&w, 1, 2);                   /* BTYPE 1: fixed codes */
    size_t i = 0;
    while (i < n) {
        unsigned best_len = 0, best_dist = 0;
        if (i + 3 <= n) {
            unsigned h = ((in[i] << 10) ^ (in[i + 1] << 5) ^ in[i + 2]) & ((1u << HB) - 1);
            int32_t c = head[h];
            int chain = CHAIN;
            size_t room = n - i;
            unsigned lim = room < max_len ? (unsigned)room : max_len;
            while (c >= 0 && chain-- > 0 && i - (size_t)c <= WIN) {
                unsigned l = 0;
                while (l < lim && in[c + l] == in[i + l]) l++;
                if (l > best_len) { best_len = l; best_dist = (unsigned)(i - (size_t)c); if (l == lim) break; }
                c = prev[c];
            }
        }
        size_t step = 1;
        if (best_len >= 3) {
            put_length(&w, best_len, max_len > 258);
            put_distance(&w, best_dist);
            step = best_len;
        } else {
            put_litlen(&w, in[i]);
        }
        for (size_t k = 0; k < step && i + k + 3 <= n; k++) {
            size_t p = i + k;
            if (step > 4096 && k > 64 && k + 64 < step) continue;   /* long runs: index the ends only */
            unsigned h = ((in[p] << 10) ^ (in[p + 1] << 5) ^ in[p + 2]) & ((1u << HB) - 1);
            prev[p] = head[h];
            head[h] = (int32_t)p;
        }
        i += step;
    }
    put_litlen(&w, 256);             /* end of block */
    if (final && w.n) put(&w, 0, 8 - w.n);
    free(head); free(prev);
    if (w.overflow) return -2;
    /* a non-final block leaves its last partial byte to the caller through *out and the return value
     * in bits is not needed by the tests: they only ever chain whole calls with final = 1 */
    return (long)w.pos;
}
