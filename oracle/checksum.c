/*
 * checksum.c -- oracle (test infrastructure): adler32 / crc32 restated from the reference,
 * plus *_combine with C zlib semantics (the reference has none).
 */
#include "zs_oracle.h"

/* common/adler32.ts:1-25 -- a += b ; s2 += a ; reduce mod 65521 every 2000 bytes.  The start
 * value is the one passed in (the "no buffer => 1" case is handled by the caller passing 1). */
uint32_t zo_adler32(uint32_t adler, const uint8_t* buf, size_t len) {
    const uint32_t BASE = 65521u;
    const size_t BLOCK = 2000;
    if (buf == NULL) return 1u;
    uint32_t lo = adler & 0xffffu;
    uint32_t hi = (adler >> 16) & 0xffffu;
    size_t pos = 0;
    while (len > 0) {
        size_t n = len > BLOCK ? BLOCK : len;
        len -= n;
        while (n--) {
            lo += buf[pos++];
            hi += lo;
        }
        lo %= BASE;
        hi %= BASE;
    }
    return (hi << 16) | lo;
}

/* common/crc32.ts:10-24 -- eight 256-entry tables for the reflected polynomial 0xEDB88320. */
static uint32_t crc_tab[8][256];
static int crc_tab_ready = 0;

__attribute__((constructor)) static void crc_tab_build(void) {
    for (unsigned n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xedb88320u ^ (c >> 1)) : (c >> 1);
        crc_tab[0][n] = c;
    }
    for (unsigned n = 0; n < 256; n++)
        for (int k = 1; k < 8; k++) {
            uint32_t p = crc_tab[k - 1][n];
            crc_tab[k][n] = (p >> 8) ^ crc_tab[0][p & 0xffu];
        }
    crc_tab_ready = 1;
}

/* common/crc32.ts:26-58 -- 8 bytes per iteration, tail bytes with table 0. */
uint32_t zo_crc32(uint32_t crc, const uint8_t* buf, size_t len) {
    if (buf == NULL) return 0u;
    if (!crc_tab_ready) crc_tab_build();
    uint32_t c = ~crc;
    size_t i = 0;
    if (len >= 8) {
        size_t end = len - 8;
        for (; i <= end; i += 8) {
            uint32_t a = c ^ ((uint32_t)buf[i] | ((uint32_t)buf[i + 1] << 8) |
                              ((uint32_t)buf[i + 2] << 16) | ((uint32_t)buf[i + 3] << 24));
            uint32_t b = (uint32_t)buf[i + 4] | ((uint32_t)buf[i + 5] << 8) |
                         ((uint32_t)buf[i + 6] << 16) | ((uint32_t)buf[i + 7] << 24);
            c = crc_tab[7][a & 0xff] ^ crc_tab[6][(a >> 8) & 0xff] ^ crc_tab[5][(a >> 16) & 0xff] ^
                crc_tab[4][a >> 24] ^ crc_tab[3][b & 0xff] ^ crc_tab[2][(b >> 8) & 0xff] ^
                crc_tab[1][(b >> 16) & 0xff] ^ crc_tab[0][b >> 24];
        }
    }
    for (; i < len; i++) c = (c >> 8) ^ crc_tab[0][(c ^ buf[i]) & 0xffu];
    return c ^ 0xffffffffu;
}

/* ---- combine (C zlib semantics) ------------------------------------------------------------
 * crc32(A||B) = crc32(A) * x^(8*len(B)) mod P  xor  crc32(B)   over GF(2)[x], reflected bit order.
 */
static uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
    /* reflected representation: bit 31 is x^0 */
    uint32_t p = 0;
    for (;;) {
        if (a & 0x80000000u) {
            p ^= b;
        }
        a <<= 1;
        if (a == 0) break;
        b = (b & 1u) ? ((b >> 1) ^ 0xedb88320u) : (b >> 1);
    }
    return p;
}

static uint32_t gf2_xpow8n(uint64_t nbytes) {
    /* x^(8*nbytes) mod P by square and multiply; x^1 is 0x40000000 in reflected form */
    uint32_t result = 0x80000000u; /* x^0 */
    uint32_t sq = 0x00800000u;     /* x^8 */
    while (nbytes) {
        if (nbytes & 1u) result = gf2_mulmod(result, sq);
        sq = gf2_mulmod(sq, sq);
        nbytes >>= 1;
    }
    return result;
}

uint32_t zo_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
    if (len2 == 0) return crc1;
    return gf2_mulmod(gf2_xpow8n(len2), crc1) ^ crc2;
}

uint32_t zo_adler32_combine(uint32_t adler1, uint32_t adler2, uint64_t len2) {
    const uint32_t BASE = 65521u;
    uint32_t rem = (uint32_t)(len2 % BASE);
    uint32_t a1 = adler1 & 0xffffu, b1 = (adler1 >> 16) & 0xffffu;
    uint32_t a2 = adler2 & 0xffffu, b2 = (adler2 >> 16) & 0xffffu;
    /* a = a1 + a2 - 1 ; b = b1 + b2 + len2*(a1 - 1)   (mod BASE) */
    uint32_t a = (a1 + a2 + BASE - 1u) % BASE;
    uint32_t b = (uint32_t)(((uint64_t)rem * ((a1 + BASE - 1u) % BASE) + b1 + b2) % BASE);
    return (b << 16) | a;
}
