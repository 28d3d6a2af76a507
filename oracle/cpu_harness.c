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
/* the LZ4 Frame layer of upstream (orig/lib/lz4frame.c, xxhash.c): interop oracle for frame.py */
#include "lz4frame.h"
#include "xxhash.h"
unsigned k4ref_xxh32(const uint8_t *p, size_t n, unsigned seed) { return XXH32(p, n, seed); }
/* one frame of independent 64 KiB blocks, acceleration 1; returns frame size or 0 */
size_t k4ref_frame_compress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, int blockChecksum, int contentChecksum)
{
    LZ4F_preferences_t prefs;
    memset(&prefs, 0, sizeof(prefs));
    prefs.frameInfo.blockSizeID = LZ4F_max64KB;
    prefs.frameInfo.blockMode = LZ4F_blockIndependent;
    prefs.frameInfo.contentChecksumFlag = contentChecksum ? LZ4F_contentChecksumEnabled : LZ4F_noContentChecksum;
    prefs.frameInfo.blockChecksumFlag = blockChecksum ? LZ4F_blockChecksumEnabled : LZ4F_noBlockChecksum;
    prefs.compressionLevel = 0;
    size_t r = LZ4F_compressFrame(dst, cap, src, n, &prefs);
    return LZ4F_isError(r) ? 0 : r;
}
size_t k4ref_frame_bound(size_t n)
{
    LZ4F_preferences_t prefs;
    memset(&prefs, 0, sizeof(prefs));
    prefs.frameInfo.blockSizeID = LZ4F_max64KB;
    prefs.frameInfo.contentChecksumFlag = LZ4F_contentChecksumEnabled;
    prefs.frameInfo.blockChecksumFlag = LZ4F_blockChecksumEnabled;
    return LZ4F_compressFrameBound(n, &prefs);
}
/* decodes a whole frame; returns decoded size, or (size_t)-1 on any error / trailing garbage */
size_t k4ref_frame_decompress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap)
{
    LZ4F_dctx *ctx = NULL;
    if (LZ4F_isError(LZ4F_createDecompressionContext(&ctx, LZ4F_VERSION))) return (size_t)-1;
    size_t ip = 0, op = 0, hint = 1;
    while (ip < n && hint != 0) {
        size_t in = n - ip, out = cap - op;
        hint = LZ4F_decompress(ctx, dst + op, &out, src + ip, &in, NULL);
        if (LZ4F_isError(hint)) { LZ4F_freeDecompressionContext(ctx); return (size_t)-1; }
        ip += in; op += out;
        if (in == 0 && out == 0) break;
    }
    LZ4F_freeDecompressionContext(ctx);
    return (hint == 0 && ip == n) ? op : (size_t)-1;
}
int k4ref_decompress_safe_usingDict(const uint8_t *s, int n, uint8_t *d, int cap, const uint8_t *dict, int dictSize)
{ return LZ4_decompress_safe_usingDict((const char *)s, (char *)d, n, cap, (const char *)dict, dictSize); }
int k4ref_decompress_safe_partial(const uint8_t *s, int n, uint8_t *d, int target)
{ return LZ4_decompress_safe_partial((const char *)s, (char *)d, n, target, target); }
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

/* Runs the block list over `threads` pthreads (contiguous ranges), each pinned to one CPU of the
 * process's affinity mask (round robin) so that passes are comparable; returns wall seconds. */
#include <sched.h>
double k4h_run_batch(int mode, const uint8_t *src, const int64_t *src_off, const int32_t *src_len,
                     uint8_t *dst, const int64_t *dst_off, const int32_t *dst_cap,
                     int32_t *out_len, int64_t n_blocks, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_t tid[1024];
    job_t jobs[1024];
    int cpus[1024], ncpu = 0;
    cpu_set_t mask;
    CPU_ZERO(&mask);
    if (sched_getaffinity(0, sizeof(mask), &mask) == 0)
        for (int c = 0; c < CPU_SETSIZE && ncpu < 1024; c++) if (CPU_ISSET(c, &mask)) cpus[ncpu++] = c;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) {
        jobs[t] = (job_t){ mode, src, src_off, src_len, dst, dst_off, dst_cap, out_len,
                           n_blocks * t / threads, n_blocks * (t + 1) / threads };
        if (threads == 1) { worker(&jobs[t]); continue; }
        pthread_attr_t at;
        pthread_attr_init(&at);
        if (ncpu > 0) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(cpus[t % ncpu], &one);
            pthread_attr_setaffinity_np(&at, sizeof(one), &one);
        }
        if (pthread_create(&tid[t], &at, worker, &jobs[t]) != 0) pthread_create(&tid[t], NULL, worker, &jobs[t]);
        pthread_attr_destroy(&at);
    }
    if (threads > 1) for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
