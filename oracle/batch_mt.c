/* batch_mt.c — TEST INFRASTRUCTURE: times the CPU restatement of the reference (this directory) on N host threads with no
 * interpreter in the loop.  It is the reference arm / cpu_baseline leg of bench.py and tools/bench_codecs.py: the Swift
 * reference runs one call per unit (Benchmarks.swift:136-145 times `Deflate.decompress(data:)` etc. in a loop); this driver
 * hands units to pthreads through one atomic counter and calls the same per-unit entry points the parity tests call.
 * Never linked into the product. */
#define _POSIX_C_SOURCE 200809L
#include <malloc.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <time.h>
#include "swco.h"

typedef struct {
    int codec;                     /* 0 deflate, 1 lz4 raw block, 2 bzip2 stream, 3 lzma2 raw (aux = dict byte), 4 xz, 5 gzip */
    const uint8_t *base;
    const uint64_t *off, *len;
    uint64_t n, total;             /* total = units to decode (indices wrap modulo n) */
    int aux;
    atomic_ullong next, bytes, failures;
} job_t;

static int decode_one(const job_t *j, uint64_t i, swco_buf *out) {
    const uint8_t *p = j->base + j->off[i];
    const size_t n = (size_t)j->len[i];
    uint64_t bits = 0;
    size_t used = 0;
    switch (j->codec) {
    case 0: return swco_deflate_decompress(p, n, 0, out, &bits);
    case 1: return swco_lz4_block(p, n, NULL, 0, out);
    case 2: return swco_bzip2_decompress(p, n, 0, out, &bits);
    case 3: return swco_lzma2_decompress_raw(p, n, (uint8_t)j->aux, out, &used);
    case 4: return swco_xz_unarchive(p, n, out);
    case 5: return swco_gzip_unarchive(p, n, out, &used);
    default: return -1;
    }
}

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    unsigned long long bytes = 0, fails = 0;
    for (;;) {
        const unsigned long long k = atomic_fetch_add_explicit(&j->next, 1, memory_order_relaxed);
        if (k >= j->total) break;
        swco_buf out = {0, 0, 0};
        const int st = decode_one(j, k % j->n, &out);
        if (st != 0) fails++;
        bytes += out.len;
        free(out.data);
    }
    atomic_fetch_add(&j->bytes, bytes);
    atomic_fetch_add(&j->failures, fails);
    return NULL;
}

/* Decodes `total` units (wrapping over the n given ones) on `nthreads` threads. Returns 0, or -1 if threads could not start. */
int swco_batch_mt(int codec, const uint8_t *base, const uint64_t *off, const uint64_t *len, uint64_t n, uint64_t total,
                  int aux, int nthreads, double *seconds, uint64_t *out_bytes, uint64_t *failures) {
    if (n == 0 || nthreads < 1) return -1;
    /* the restatement allocates a 2^(maxBits+1)-slot heap per Huffman tree like DecodingTree.swift:22-32 does; keep those
     * blocks in per-thread malloc arenas instead of mmap/munmap, whose address-space lock would serialise the threads */
    mallopt(M_MMAP_THRESHOLD, 16 << 20);
    mallopt(M_TRIM_THRESHOLD, 512 << 20);
    job_t j;
    j.codec = codec; j.base = base; j.off = off; j.len = len; j.n = n; j.total = total; j.aux = aux;
    atomic_init(&j.next, 0); atomic_init(&j.bytes, 0); atomic_init(&j.failures, 0);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    if (!th) return -1;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int started = 0;
    for (; started < nthreads; started++)
        if (pthread_create(&th[started], NULL, worker, &j) != 0) break;
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    if (started == 0) return -1;
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    *out_bytes = atomic_load(&j.bytes);
    *failures = atomic_load(&j.failures);
    return 0;
}
