/*
 * cpu_harness.c -- multi-threaded CPU timing harness around a block codec
 * (TEST / BASELINE INFRASTRUCTURE ONLY; used by bench.py's cpu_baseline and
 * --impl reference legs).  Built twice by oracle/Makefile:
 *   -DK4H_USE_REF : calls LZ4_compress_fast / LZ4_decompress_safe from the reference's
 *                   own upstream C source (orig/lib/lz4.c, compiled where it lies)
 *                   -> oracle/_ref/libk4ref.so          (cpu_baseline.kind = "reference")
 *   default       : calls the restatement in k4lz4_oracle.c
 *                   -> oracle/_build/libk4oracle.so     (cpu_baseline.kind = "port")
 * The call shapes are the ones LZ4Codec.Encode/Decode make (LZ4Codec.cs:40-52,104-115):
 * per block, fresh state, acceleration 1.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef K4H_USE_REF
#include "lz4.h"
static int enc_one(const uint8_t *s, int n, uint8_t *d, int cap)
{ int r = LZ4_compress_fast((const char *)s, (char *)d, n, cap, 1); return r <= 0 ? -1 : r; }
static int dec_one(const uint8_t *s, int n, uint8_t *d, int cap)
{ int r = LZ4_decompress_safe((const char *)s, (char *)d, n, cap); return r <= 0 ? -1 : r; }
int k4h_is_reference(void) { return 1; }
/* thin exports so tests can call the reference engine directly */
int k4ref_compress_fast(const uint8_t *s, int n, uint8_t *d, int cap)
{ return LZ4_compress_fast((const char *)s, (char *)d, n, cap, 1); }
int k4ref_decompress_safe(const uint8_t *s, int n, uint8_t *d, int cap)
{ return LZ4_decompress_safe((const char *)s, (char *)d, n, cap); }
int k4ref_version(void) { return LZ4_versionNumber(); }
/* the reference's own synthetic-data generator (orig/programs/datagen.c:156-162), SURVEY 8(d) */
#include <stddef.h>
void RDG_genBuffer(void *buffer, size_t size, double matchProba, double litProba, unsigned seed);
int k4ref_datagen(uint8_t *buf, size_t size, double matchProba, double litProba, unsigned seed)
{ if (!buf || matchProba >= 1.0) return -1; RDG_genBuffer(buf, size, matchProba, litProba, seed); return 0; }
#else
#include "k4lz4_oracle.h"
static int enc_one(const uint8_t *s, int n, uint8_t *d, int cap)
{ return k4o_codec_encode(s, n, d, cap, 0, 0); }
static int dec_one(const uint8_t *s, int n, uint8_t *d, int cap)
{ return k4o_codec_decode(s, n, d, cap); }
int k4h_is_reference(void) { return 0; }
#endif

typedef struct {
    int mode;                 /* 0 = encode, 1 = decode */
    const uint8_t *src; const int64_t *src_off; const int32_t *src_len;
    uint8_t *dst; const int64_t *dst_off; const int32_t *dst_cap;
    int32_t *out_len;
    int64_t lo, hi;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    for (int64_t i = j->lo; i < j->hi; i++) {
        const uint8_t *s = j->src + j->src_off[i];
        uint8_t *d = j->dst + j->dst_off[i];
        int32_t n = j->src_len[i];
        int r;
        if (n <= 0) r = 0;
        else r = j->mode == 0 ? enc_one(s, n, d, j->dst_cap[i]) : dec_one(s, n, d, j->dst_cap[i]);
        j->out_len[i] = r;
    }
    return NULL;
}

/* Runs the block list over `threads` pthreads (contiguous ranges); returns wall seconds. */
double k4h_run_batch(int mode, const uint8_t *src, const int64_t *src_off, const int32_t *src_len,
                     uint8_t *dst, const int64_t *dst_off, const int32_t *dst_cap,
                     int32_t *out_len, int64_t n_blocks, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_t tid[1024];
    job_t jobs[1024];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) {
        jobs[t] = (job_t){ mode, src, src_off, src_len, dst, dst_off, dst_cap, out_len,
                           n_blocks * t / threads, n_blocks * (t + 1) / threads };
        if (threads == 1) worker(&jobs[t]);
        else pthread_create(&tid[t], NULL, worker, &jobs[t]);
    }
    if (threads > 1) for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
