/*
 * zs_oracle.h -- CPU restatement of the zlib-streams-ts hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and only
 * as the checker or the CPU baseline.  The product path is libzsgpu.so (CUDA, sm_100a) and has
 * no CPU fallback.
 *
 * Each function cites the reference file:line (relative to the zlib-streams-ts tree) whose
 * behaviour it restates.  The reference itself is TypeScript and cannot be executed in this image
 * (no node/tsx); the restatement is pinned by (1) the known-answer vectors of the reference's own
 * test-suite, (2) the deflate64 fixtures of test/data, (3) C zlib 1.3, which the reference is a
 * line-by-line port of and which 27 of its own test files use as the cross-oracle.
 */
#ifndef ZS_ORACLE_H
#define ZS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* return codes, common/constants.ts:21-29 */
#define ZO_OK 0
#define ZO_STREAM_END 1
#define ZO_NEED_DICT 2
#define ZO_ERRNO (-1)
#define ZO_STREAM_ERROR (-2)
#define ZO_DATA_ERROR (-3)
#define ZO_MEM_ERROR (-4)
#define ZO_BUF_ERROR (-5)

/* flush values, common/constants.ts:13-19 */
#define ZO_NO_FLUSH 0
#define ZO_PARTIAL_FLUSH 1
#define ZO_SYNC_FLUSH 2
#define ZO_FULL_FLUSH 3
#define ZO_FINISH 4
#define ZO_BLOCK 5
#define ZO_TREES 6

/* ---- checksums: common/adler32.ts, common/crc32.ts -------------------------------------- */
uint32_t zo_adler32(uint32_t adler, const uint8_t* buf, size_t len);
uint32_t zo_crc32(uint32_t crc, const uint8_t* buf, size_t len);
/* No reference analogue (the TS library has no *_combine); semantics are C zlib's. */
uint32_t zo_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);
uint32_t zo_adler32_combine(uint32_t adler1, uint32_t adler2, uint64_t len2);

/* ---- inflate: inflate/inflate.ts, inftrees.ts, inffast.ts ------------------------------- */

/* inflate_table, inftrees.ts:62.  type 0 CODES, 1 LENS, 2 DISTS.  `table` receives packed
 * entries op<<24|bits<<16|val starting at table[*index]; *index advances by the space used. */
int zo_inflate_table(int type, const uint16_t* lens, unsigned codes, uint32_t* table,
                     unsigned* bits, uint16_t* work, unsigned* index, int deflate64);

typedef struct zo_inflate_stream zo_inflate_stream;

zo_inflate_stream* zo_inflate_new(void);                       /* createInflateStream, inflate.ts:68 */
void zo_inflate_free(zo_inflate_stream* s);
int zo_inflate_init2(zo_inflate_stream* s, int window_bits);    /* inflateInit2_, inflate.ts:174 */
int zo_inflate_reset(zo_inflate_stream* s);                     /* inflateReset, inflate.ts:124 */
int zo_inflate_set_dictionary(zo_inflate_stream* s, const uint8_t* dict, size_t n); /* :1220 */
int zo_inflate(zo_inflate_stream* s, int flush);                /* inflate, inflate.ts:332 */
int zo_inflate_end(zo_inflate_stream* s);                       /* inflateEnd, inflate.ts:1187 */

/* stream counters (the Stream carrier of common/types.ts:1-15) */
void zo_inflate_set_input(zo_inflate_stream* s, const uint8_t* p, size_t n);
void zo_inflate_set_output(zo_inflate_stream* s, uint8_t* p, size_t n);
size_t zo_inflate_avail_in(const zo_inflate_stream* s);
size_t zo_inflate_avail_out(const zo_inflate_stream* s);
uint64_t zo_inflate_total_in(const zo_inflate_stream* s);
uint64_t zo_inflate_total_out(const zo_inflate_stream* s);
uint32_t zo_inflate_adler(const zo_inflate_stream* s);
const char* zo_inflate_msg(const zo_inflate_stream* s);
int zo_inflate_mode(const zo_inflate_stream* s);

/* One-shot helper: inflate(strm, Z_FINISH) of a whole buffer.  Returns the inflate() return
 * code; *out_len / *in_used report total_out / total_in; *check the running adler/crc. */
int zo_inflate_oneshot(const uint8_t* in, size_t in_len, int window_bits, const uint8_t* dict,
                       size_t dict_len, uint8_t* out, size_t out_cap, size_t* out_len,
                       size_t* in_used, uint32_t* check);

/* Test aid: how many stored / fixed / dynamic blocks the last zo_inflate_oneshot call walked through
 * (counts[3]); not thread safe. */
void zo_inflate_last_blocks(uint32_t* counts);

/* Test aid: raw deflate64 encoder (one fixed-Huffman block, greedy matching in a 64 KiB window); the
 * reference has none.  max_len 258 stays within length codes <= 284, 65538 allows code 285.  Returns
 * the bytes written, -1 out of memory, -2 output too small.  See deflate64_enc.c. */
long zo_deflate64_encode(const uint8_t* in, size_t n, unsigned max_len, int final, uint8_t* out, size_t cap);

/* ---- deflate: deflate/deflate.ts, trees.ts, deflate/utils.ts ---------------------------- */

/* Upper bound of deflateBound for windowBits 15 / memLevel 8, deflate.ts:615-674.
 * wrap: 0 raw, 1 zlib, 2 gzip. */
size_t zo_deflate_bound(size_t source_len, int wrap);

/* One-shot deflate of `in[0..in_len)`: deflateInit2_(level, 8, wbits, 8, 0) [+ deflateSetDictionary
 * (dict) when dict_len > 0] then deflate(flush) with the whole input, flush = ZO_FINISH or
 * ZO_SYNC_FLUSH (to produce a non-final, byte-aligned chunk).  wrap: 0 raw, 1 zlib, 2 gzip.
 * Returns the number of bytes written, or a negative ZO_* code. */
int64_t zo_deflate_oneshot(const uint8_t* in, size_t in_len, int level, int wrap,
                           const uint8_t* dict, size_t dict_len, int flush, uint8_t* out,
                           size_t out_cap);

/* The same with every level and strategy of deflateInit2_ (deflate.ts:253-297): level 0 is
 * deflate_stored (deflate.ts:1140-1279), ZO_FILTERED changes deflate_slow's short-match rule
 * (:1386-1392), ZO_HUFFMAN_ONLY is deflate_huff (:1525-1560), ZO_FIXED forces static trees
 * (trees.ts:567).  ZO_RLE is C zlib's deflate_rle -- the algorithm the port intends and the engine
 * implements; with rle_like_reference != 0 it is what the reference really emits for Z_RLE: its
 * scan compares a byte with an index (deflate.ts:1467-1469), never finds a run, and degenerates to
 * deflate_huff's output. */
#define ZO_DEFAULT_STRATEGY 0
#define ZO_FILTERED 1
#define ZO_HUFFMAN_ONLY 2
#define ZO_RLE 3
#define ZO_FIXED 4
int64_t zo_deflate_oneshot2(const uint8_t* in, size_t in_len, int level, int strategy,
                            int rle_like_reference, int wrap, const uint8_t* dict, size_t dict_len,
                            int flush, uint8_t* out, size_t out_cap);

/* build_tree + gen_bitlen + gen_codes (trees.ts:54-76,187-316) for one tree, exposed so the GPU
 * code-length construction can be compared entry by entry.  kind: 0 literal/length (286 symbols,
 * limit 15), 1 distance (30, limit 15), 2 bit-length (19, limit 7).  freq is modified like the
 * reference does (forcing two codes).  Returns max_code. */
int zo_build_tree(int kind, uint16_t* freq, uint16_t* len_out, uint16_t* code_out,
                  uint32_t* opt_len, uint32_t* static_len);

#ifdef __cplusplus
}
#endif
#endif
