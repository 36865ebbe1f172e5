/*
 * bench_driver.c -- oracle (test infrastructure): multi-threaded driver that runs the CPU
 * restatement over many independent chunks / records, used ONLY by bench.py's cpu_baseline leg and
 * by `bench.py --impl reference` (one independent stream per host thread, like one reference
 * stream per worker_thread).  pthreads + an atomic work counter.
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "zs_oracle.h"

int zo_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct job job_t;
typedef void (*item_fn)(job_t*, size_t item, void* tls);
struct job {
    atomic_size_t next;
    size_t n_items, grain;
    item_fn fn;
    /* deflate */
    const uint8_t* in; size_t in_len, chunk; int level, wrap, prime;
    _Atomic int64_t total; atomic_int bad;
    /* inflate */
    const uint64_t* in_off; int window_bits; uint8_t* out; const uint64_t* out_off;
    uint64_t* out_len; uint32_t* checks; int32_t* status;
    /* checksum */
    int kind; uint32_t* part;
};

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    size_t cap = zo_deflate_bound(j->chunk ? j->chunk : 1, 2) + 64;
    void* tls = malloc(cap);
    for (;;) {
        size_t b = atomic_fetch_add(&j->next, j->grain);
        if (b >= j->n_items) break;
        size_t e = b + j->grain < j->n_items ? b + j->grain : j->n_items;
        for (size_t i = b; i < e; i++) j->fn(j, i, tls);
    }
    free(tls);
    return NULL;
}

static void run_job(job_t* j, int threads) {
    if (threads <= 0) threads = zo_max_threads();
    if ((size_t)threads > j->n_items) threads = j->n_items ? (int)j->n_items : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    atomic_store(&j->next, 0);
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, worker, j);
    worker(j);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
}

static void deflate_item(job_t* j, size_t i, void* tls) {
    size_t off = i * j->chunk;
    size_t n = j->in_len - off < j->chunk ? j->in_len - off : j->chunk;
    size_t dl = j->prime ? (off < 32768 ? off : 32768) : 0;
    int last = i + 1 == j->n_items;
    int solo = last || !j->prime;
    size_t cap = zo_deflate_bound(j->chunk, 2) + 64;
    int64_t r = zo_deflate_oneshot(j->in + off, n, j->level, solo ? j->wrap : 0, dl ? j->in + off - dl : NULL, dl,
                                   solo ? ZO_FINISH : ZO_SYNC_FLUSH, (uint8_t*)tls, cap);
    if (r < 0) atomic_store(&j->bad, 1); else atomic_fetch_add(&j->total, r);
}

/* Deflate the chunks of `chunk` bytes of `in` (last one may be short), each primed with the up to
 * 32 KiB that precede it when `prime` is set (deflateSetDictionary + Z_SYNC_FLUSH for all but the
 * last chunk); returns total compressed bytes or -1. */
int64_t zo_deflate_chunks_mt(const uint8_t* in, size_t in_len, size_t chunk, int level, int wrap, int prime,
                             int threads) {
    job_t j;
    memset(&j, 0, sizeof j);
    j.n_items = (in_len + chunk - 1) / chunk; j.grain = 1; j.fn = deflate_item;
    j.in = in; j.in_len = in_len; j.chunk = chunk; j.level = level; j.wrap = wrap; j.prime = prime;
    run_job(&j, threads);
    return atomic_load(&j.bad) ? -1 : atomic_load(&j.total);
}

static void deflate_store_item(job_t* j, size_t i, void* tls) {
    (void)tls;
    size_t off = i * j->chunk;
    size_t n = j->in_len - off < j->chunk ? j->in_len - off : j->chunk;
    const size_t slot = (size_t)j->out_off[0];
    int64_t r = zo_deflate_oneshot(j->in + off, n, j->level, j->wrap, NULL, 0, ZO_FINISH, j->out + i * slot, slot);
    if (r < 0) { atomic_store(&j->bad, 1); r = 0; }
    j->out_len[i] = (uint64_t)r;
}

/* Deflate every `chunk`-byte record of `in` as an independent stream (wrap 0 raw, 1 zlib, 2 gzip) and keep
 * the streams: packed back to back into `out` (capacity n_records * slot, slot >= zo_deflate_bound(chunk, wrap)),
 * offsets to out_off[n_records + 1].  This is how bench.py makes the reference-produced record set of
 * BASELINE configs[3].  Returns the total size or -1. */
int64_t zo_deflate_records_mt(const uint8_t* in, size_t in_len, size_t chunk, int level, int wrap, uint8_t* out,
                              size_t slot, uint64_t* out_off, int threads) {
    job_t j;
    memset(&j, 0, sizeof j);
    j.n_items = (in_len + chunk - 1) / chunk; j.grain = 16; j.fn = deflate_store_item;
    j.in = in; j.in_len = in_len; j.chunk = chunk; j.level = level; j.wrap = wrap;
    uint64_t slot64 = slot;
    uint64_t* lens = (uint64_t*)malloc((j.n_items + 1) * sizeof(uint64_t));
    j.out = out; j.out_off = &slot64; j.out_len = lens;
    run_job(&j, threads);
    if (atomic_load(&j.bad)) { free(lens); return -1; }
    uint64_t pos = 0;
    for (size_t i = 0; i < j.n_items; i++) {
        memmove(out + pos, out + i * slot, (size_t)lens[i]);
        out_off[i] = pos;
        pos += lens[i];
    }
    out_off[j.n_items] = pos;
    free(lens);
    return (int64_t)pos;
}

static void inflate_item(job_t* j, size_t i, void* tls) {
    (void)tls;
    size_t ol = 0, used = 0;
    uint32_t ck = 0;
    int r = zo_inflate_oneshot(j->in + j->in_off[i], (size_t)(j->in_off[i + 1] - j->in_off[i]), j->window_bits,
                               NULL, 0, j->out + j->out_off[i], (size_t)(j->out_off[i + 1] - j->out_off[i]), &ol,
                               &used, &ck);
    if (j->out_len) j->out_len[i] = ol;
    if (j->checks) j->checks[i] = ck;
    if (j->status) j->status[i] = r;
    if (r != ZO_STREAM_END) atomic_fetch_add(&j->total, 1);
}

/* Inflate n independent streams: in_off[n+1] byte offsets into `in`, out_off[n+1] into `out`.
 * Returns the number of streams that did not end with ZO_STREAM_END. */
int64_t zo_inflate_batch_mt(const uint8_t* in, const uint64_t* in_off, size_t n, int window_bits, uint8_t* out,
                            const uint64_t* out_off, uint64_t* out_len, uint32_t* checks, int32_t* status,
                            int threads) {
    job_t j;
    memset(&j, 0, sizeof j);
    j.n_items = n; j.grain = 32; j.fn = inflate_item;
    j.in = in; j.in_off = in_off; j.window_bits = window_bits; j.out = out; j.out_off = out_off;
    j.out_len = out_len; j.checks = checks; j.status = status;
    run_job(&j, threads);
    return atomic_load(&j.total);
}

static void checksum_item(job_t* j, size_t i, void* tls) {
    (void)tls;
    size_t off = i * j->chunk;
    size_t n = j->in_len - off < j->chunk ? j->in_len - off : j->chunk;
    j->part[i] = j->kind ? zo_crc32(0u, j->in + off, n) : zo_adler32(1u, j->in + off, n);
}

/* checksum of the chunks, folded with *_combine; kind 0 adler32, 1 crc32 */
uint32_t zo_checksum_mt(const uint8_t* in, size_t in_len, size_t chunk, int kind, int threads) {
    job_t j;
    memset(&j, 0, sizeof j);
    j.n_items = (in_len + chunk - 1) / chunk;
    if (j.n_items == 0) return kind ? 0u : 1u;
    j.grain = 1; j.fn = checksum_item; j.in = in; j.in_len = in_len; j.chunk = chunk; j.kind = kind;
    j.part = (uint32_t*)malloc(j.n_items * sizeof(uint32_t));
    run_job(&j, threads);
    uint32_t acc = j.part[0];
    for (size_t i = 1; i < j.n_items; i++) {
        size_t off = i * chunk;
        size_t n = in_len - off < chunk ? in_len - off : chunk;
        acc = kind ? zo_crc32_combine(acc, j.part[i], n) : zo_adler32_combine(acc, j.part[i], n);
    }
    free(j.part);
    return acc;
}
