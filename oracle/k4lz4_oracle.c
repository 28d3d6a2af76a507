/*
 * k4lz4_oracle.c -- CPU restatement of the K4os.Compression.LZ4 block hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it.  The product library
 * (libk4lz4.so) never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  This restatement is checked (tests/test_oracle.py) against
 *   (1) the reference's own golden decode vector assets/issue64 (Issue64.cs:16-55),
 *       committed as tests/golden/issue64_block0.*;
 *   (2) the reference's own oracle -- native lz4 (orig/lib/lz4.c) compiled as-is into
 *       oracle/_ref/libk4ref.so by oracle/Makefile -- the code that generated the
 *       reference's ChecksumBlockTests golden rows (playground/SharedSources/app.cpp:94-97);
 *   (3) golden encode rows produced by (2) in this container and committed under
 *       tests/golden/encode_rows.json (generator: tests/golden/make_golden.py).
 *
 * All file:line citations are relative to /root/reference/src/K4os.Compression.LZ4/.
 * The code below is written from the behaviour of those lines;
 * it shares no text with them: indices instead of pointers, one specialised loop per
 * table type instead of the reference's directive-switch, byte loops instead of wild copies.
 */
#include <stdint.h>
#include <string.h>
#include <stddef.h>

#include "k4lz4_oracle.h"

/* ---- constants: Engine/LL.types.cs:50-78 ------------------------------------------- */
enum {
    K_MINMATCH = 4,
    K_LASTLITERALS = 5,
    K_MFLIMIT = 12,
    K_MINLENGTH = 13,           /* LZ4_minLength = MFLIMIT + 1 */
    K_64KLIMIT = 65536 + 11,    /* LZ4_64Klimit */
    K_MAXDIST = 65535,          /* LZ4_DISTANCE_MAX */
    K_SKIPTRIGGER = 6,
    K_MAX_INPUT = 0x7E000000
};

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* Engine/LL.tools.cs:38-40 (LZ4_compressBound) == LZ4Codec.MaximumOutputSize (LZ4Codec.cs:30-31) */
int k4o_max_output_size(int n)
{
    return n > K_MAX_INPUT ? 0 : n + n / 255 + 16;
}

/* Engine/LL.tools.cs:46-51 (LZ4_hash4): hashLog 13 for the u16 table, 12 otherwise. */
static inline uint32_t hash4(uint32_t v, int log) { return (v * 2654435761u) >> (32 - log); }
/* Engine/LL.tools.cs:53-58 (LZ4_hash5), little-endian form. */
static inline uint32_t hash5(uint64_t v, int log)
{
    return (uint32_t)(((v << 24) * 889523592379ull) >> (64 - log));
}

/* Engine/x64/LL64.tools.cs:87-133 (LZ4_count).  Net effect: number of equal bytes of
 * src[a..] vs src[b..] with a < limit.  The 8/4/2/1 stepping of the reference yields the
 * plain common-prefix length, which is what is computed here. */
static inline uint32_t common_len(const uint8_t *s, uint32_t a, uint32_t b, uint32_t limit)
{
    uint32_t start = a;
    while (a < limit && s[a] == s[b]) { a++; b++; }
    return a - start;
}

/* literal-length / last-run emitter shared by LL64.fast.cs:261-268 and :488-498 */
static inline uint32_t put_run_token(uint8_t *dst, uint32_t op, uint32_t run, uint32_t *tokpos)
{
    *tokpos = op;
    if (run >= 15) {
        uint32_t rest = run - 15;
        dst[op++] = 0xF0;
        for (; rest >= 255; rest -= 255) dst[op++] = 255;
        dst[op++] = (uint8_t)rest;
    } else {
        dst[op++] = (uint8_t)(run << 4);
    }
    return op;
}

/*
 * LZ4_compress_generic specialised as LZ4_compress_fast_extState does it
 * (Engine/x64/LL64.fast.cs:517-568): noDict, noDictIssue, acceleration 1,
 *   table  = byU16 (8192 x u16, hash4/13 bits)       when n < 65547
 *          = byU32 (4096 x u32, hash5/12 bits on X64; hash4/12 bits on X32/Enforce32)
 *   output = notLimited when cap >= compressBound(n), else limitedOutput.
 * Returns the engine's value: bytes written, 0 when it does not fit.
 * Follows Engine/x64/LL64.fast.cs:35-513 step by step (step numbers = SURVEY.md App. A).
 */
int k4o_compress_fast(const uint8_t *src, int n, uint8_t *dst, int cap, int enforce32)
{
    /* LZ4_stream_t hash table, zeroed: LL.tools.cs:235-239, LL.types.cs:30-39 */
    static _Thread_local uint32_t table32[4096];
    uint16_t *table16 = (uint16_t *)table32;
    memset(table32, 0, sizeof(table32));

    if ((uint32_t)n > (uint32_t)K_MAX_INPUT) return 0;                 /* :90 */

    const int by16 = n < K_64KLIMIT;                                   /* :526,548 */
    const int limited = !(cap >= k4o_max_output_size(n));              /* :524 */
    const int64_t olimit = cap;
    const int use5 = !by16 && !enforce32;                              /* LL64.tools.cs:135-143 */
    const int hlog = by16 ? 13 : 12;

#define HASH_AT(p) (use5 ? hash5(rd64(src + (p)), hlog) : hash4(rd32(src + (p)), hlog))
#define TGET(h)    (by16 ? (uint32_t)table16[h] : table32[h])
#define TPUT(h, v) do { if (by16) table16[h] = (uint16_t)(v); else table32[h] = (uint32_t)(v); } while (0)

    uint32_t ip = 0, anchor = 0;
    int64_t op = 0;
    const uint32_t un = (uint32_t)n;

    if (n >= K_MINLENGTH) {                                            /* :117 */
        const uint32_t mfl1 = un - K_MFLIMIT + 1;                      /* mflimitPlusOne :70 */
        const uint32_t mlim = un - K_LASTLITERALS;                     /* matchlimit :71 */
        uint32_t fh, m, tokpos;

        TPUT(HASH_AT(0), 0);                                           /* :120 */
        ip = 1;
        fh = HASH_AT(1);                                               /* :122 */

        for (;;) {
            /* step 3 -- search, :158-234 */
            {
                uint32_t fwd = ip, step = 1, cnt = 1u << K_SKIPTRIGGER;
                for (;;) {
                    uint32_t hh = fh, cur = fwd;
                    m = TGET(hh);
                    ip = fwd;
                    fwd += step;
                    step = cnt++ >> K_SKIPTRIGGER;
                    if (fwd > mfl1) goto last_literals;                /* :172 */
                    fh = HASH_AT(fwd);                                 /* :212 */
                    TPUT(hh, cur);                                     /* :213 */
                    if (!by16 && m + K_MAXDIST < cur) continue;        /* :219-224 */
                    if (rd32(src + m) == rd32(src + ip)) break;        /* :228 */
                }
            }
            /* step 4 -- catch-up, :237-242 (lowLimit == source) */
            while (ip > anchor && m > 0 && src[ip - 1] == src[m - 1]) { ip--; m--; }

            /* step 5 -- literals, :244-272 */
            {
                uint32_t lit = ip - anchor;
                if (limited && op + 1 + lit + 8 + lit / 255 > olimit) return 0;   /* :246-251 */
                op = put_run_token(dst, (uint32_t)op, lit, &tokpos);
                memcpy(dst + op, src + anchor, lit);                   /* WildCopy8 net effect */
                op += lit;
            }

        next_match:
            /* step 6 -- offset + match length, :291-382 */
            dst[op] = (uint8_t)(ip - m);
            dst[op + 1] = (uint8_t)((ip - m) >> 8);
            op += 2;
            {
                uint32_t mc = common_len(src, ip + K_MINMATCH, m + K_MINMATCH, mlim);   /* :328 */
                ip += mc + K_MINMATCH;
                if (limited && op + 6 + (mc + 240) / 255 > olimit) return 0;           /* :332-362 */
                if (mc >= 15) {
                    dst[tokpos] += 15;
                    mc -= 15;
                    /* :369-378 writes 0xFF words then lands on op + mc/255; net effect: */
                    memset(dst + op, 0xFF, mc / 255);
                    op += mc / 255;
                    dst[op++] = (uint8_t)(mc % 255);
                } else {
                    dst[tokpos] += (uint8_t)mc;
                }
            }
            anchor = ip;                                               /* :388 */
            if (ip >= mfl1) break;                                     /* :391 */

            /* step 8 -- post-match insert and probe, :394-466 */
            TPUT(HASH_AT(ip - 2), ip - 2);
            {
                uint32_t hh = HASH_AT(ip);
                m = TGET(hh);
                TPUT(hh, ip);
                if ((by16 || m + K_MAXDIST >= ip) && rd32(src + m) == rd32(src + ip)) {
                    tokpos = (uint32_t)op;
                    dst[op++] = 0;
                    goto next_match;
                }
            }
            fh = HASH_AT(++ip);                                        /* :466 */
        }
    }

last_literals:
    /* step 9 -- :469-503 */
    {
        uint32_t run = un - anchor, tokpos;
        if (limited && op + run + 1 + (run + 255 - 15) / 255 > olimit) return 0;
        op = put_run_token(dst, (uint32_t)op, run, &tokpos);
        memcpy(dst + op, src + anchor, run);
        op += run;
    }
    return (int)op;
#undef HASH_AT
#undef TGET
#undef TPUT
}

/*
 * LZ4_decompress_safe (Engine/x64/LL64.dec.cs:469-477) == LZ4_decompress_generic with
 * endOnInputSize, full, noDict, lowPrefix = dst, dictSize = 0 (=> checkOffset).
 * Follows Engine/x64/LL64.dec.cs:124-467 check by check, in the same order, so that the
 * accept/reject decision and the returned value equal the reference's for EVERY input,
 * well-formed or not.  Bytes are produced with sequential byte copies (the LZ77
 * meaning of the reference's Copy8/16/18/WildCopy8 sequences).
 * One deliberate, documented difference: a match with offset 0 (accepted by the
 * reference, never produced by an encoder) reads "whatever dst held before"; here it
 * yields zero bytes.  Only the returned length is contractual for such streams.
 * Returns bytes written (>= 0) or a negative error position like the reference.
 */
int k4o_decompress_safe(const uint8_t *src, int n, uint8_t *dst, int cap)
{
    if (src == NULL) return -1;                                        /* :136 */
    int64_t ip = 0, op = 0;
    const int64_t iend = n, oend = cap;
    const int64_t shortiend = iend - 14 - 2;                           /* :152 */
    const int64_t shortoend = oend - 14 - 18;                          /* :153 */

    if (cap == 0) return (n == 1 && src[0] == 0) ? 0 : -1;             /* :162-168 */
    if (n == 0) return -1;                                             /* :172 */

    for (;;) {
        uint32_t token = src[ip++];                                    /* :177 */
        int64_t length = token >> 4;
        int64_t offset, match, cpy;

        if (length != 15 && ip < shortiend && op <= shortoend) {       /* :191-193 */
            memcpy(dst + op, src + ip, (size_t)length);                /* Copy16, net */
            op += length; ip += length;
            length = token & 15;                                       /* :204 */
            offset = src[ip] | (src[ip + 1] << 8); ip += 2;
            match = op - offset;
            if (length != 15 && offset >= 8 && match >= 0) {           /* :211-213 */
                length += K_MINMATCH;
                for (int64_t i = 0; i < length; i++) dst[op + i] = dst[match + i];
                op += length;
                continue;
            }
            goto copy_match;                                           /* :224 */
        }

        if (length == 15) {                                            /* :228-243 */
            /* LZ4_readVLE(lencheck = iend-15, loop_check, initial_check) LL.tools.cs:165-193.
             * Only initial_error is fatal here (:232); a loop_error just stops the sum. */
            if (ip >= iend - 15) goto output_error;
            for (;;) {
                uint32_t s = src[ip++];
                length += s;
                if (ip >= iend - 15) break;
                if (s != 255) break;
            }
        }

        cpy = op + length;                                             /* :246 */
        if (cpy > oend - K_MFLIMIT || ip + length > iend - (2 + 1 + K_LASTLITERALS)) {
            if (ip + length != iend || cpy > oend) goto output_error;  /* :291-294 */
            memmove(dst + op, src + ip, (size_t)length);               /* :297 */
            ip += length; op += length;
            break;                                                     /* :304-307 */
        }
        memcpy(dst + op, src + ip, (size_t)length);                    /* :311 */
        ip += length; op = cpy;

        offset = src[ip] | (src[ip + 1] << 8); ip += 2;                /* :318-320 */
        match = op - offset;
        length = token & 15;                                           /* :323 */

    copy_match:
        if (length == 15) {                                            /* :326-334 */
            /* LZ4_readVLE(lencheck = iend-4, loop_check, no initial_check): any overrun fatal */
            for (;;) {
                uint32_t s = src[ip++];
                length += s;
                if (ip >= iend - K_LASTLITERALS + 1) goto output_error;
                if (s != 255) break;
            }
        }
        length += K_MINMATCH;                                          /* :336 */
        if (match < 0) goto output_error;                              /* :338 */
        cpy = op + length;                                             /* :383 */
        if (cpy > oend - 12 && cpy > oend - K_LASTLITERALS) goto output_error;   /* :427-433 */
        if (offset == 0) {
            memset(dst + op, 0, (size_t)length);                       /* see header note */
        } else {
            for (int64_t i = 0; i < length; i++) dst[op + i] = dst[match + i];
        }
        op = cpy;                                                      /* :450 */
    }
    return (int)op;                                                    /* :454-457 */

output_error:
    return (int)(-ip) - 1;                                             /* :465 */
}

/*
 * The general form of the same loop: LZ4_decompress_generic with endOnInputSize and
 *   - an external dictionary (usingExtDict, lowPrefix = dst): LZ4_decompress_safe_forceExtDict,
 *     Engine/x64/LL64.dec.cs:510-521, reached from LZ4_decompress_safe_usingDict :523-546.  The
 *     prefix variants of :529-541 (dictionary adjacent to dst) read the same bytes through plain
 *     pointers and reject the same offsets (match + dictSize < lowPrefix, :338), so one
 *     restatement covers all three; a dictionary of >= 64 KiB switches the offset test off
 *     (checkOffset, :147) -- every 16-bit offset then lands inside it.
 *   - earlyEnd = partial: LZ4_decompress_safe_partial :548-556 (LLxx.cs:29-39 passes
 *     targetOutputSize == dstCapacity == the target length), paths :256-280, :301-307, :387-406.
 * Ext-dict matches follow :341-378 (own length test, then dictionary part + block part); their
 * bytes are those of the virtual window [dict | dst].
 */
static int decompress_general(const uint8_t *src, int n, uint8_t *dst, int outputSize, int partial,
                              const uint8_t *dict, int dictSize)
{
    if (src == NULL) return -1;                                        /* :136 */
    int64_t ip = 0, op = 0;
    const int64_t iend = n, oend = outputSize;
    const int64_t shortiend = iend - 14 - 2, shortoend = oend - 14 - 18;
    const int checkOffset = dictSize < 65536;                          /* :147 */
    const int extDict = dict != NULL && dictSize > 0;

    if (outputSize == 0) {                                             /* :162-168 */
        if (partial) return 0;
        return (n == 1 && src[0] == 0) ? 0 : -1;
    }
    if (n == 0) return -1;                                             /* :172 */

    for (;;) {
        uint32_t token = src[ip++];
        int64_t length = token >> 4;
        int64_t offset, match, cpy;

        if (length != 15 && ip < shortiend && op <= shortoend) {       /* :191-193 */
            memcpy(dst + op, src + ip, (size_t)length);
            op += length; ip += length;
            length = token & 15;
            offset = src[ip] | (src[ip + 1] << 8); ip += 2;
            match = op - offset;
            if (length != 15 && offset >= 8 && match >= 0) {           /* :211-213 (match >= lowPrefix) */
                length += K_MINMATCH;
                for (int64_t i = 0; i < length; i++) dst[op + i] = dst[match + i];
                op += length;
                continue;
            }
            goto copy_match;
        }

        if (length == 15) {                                            /* :228-243 */
            if (ip >= iend - 15) goto output_error;
            for (;;) {
                uint32_t s = src[ip++];
                length += s;
                if (ip >= iend - 15) break;
                if (s != 255) break;
            }
        }

        cpy = op + length;                                             /* :246 */
        if (cpy > oend - K_MFLIMIT || ip + length > iend - (2 + 1 + K_LASTLITERALS)) {
            if (partial) {                                             /* :256-280 */
                if (ip + length > iend - (2 + 1 + K_LASTLITERALS) && ip + length != iend) goto output_error;
                if (cpy > oend) { cpy = oend; length = oend - op; }
            } else {
                if (ip + length != iend || cpy > oend) goto output_error;   /* :291-294 */
            }
            memmove(dst + op, src + ip, (size_t)length);               /* :297 */
            ip += length; op += length;
            if (!partial || cpy == oend || ip == iend) break;          /* :304-307 */
        } else {
            memcpy(dst + op, src + ip, (size_t)length);                /* :311 */
            ip += length; op = cpy;
        }

        offset = src[ip] | (src[ip + 1] << 8); ip += 2;                /* :318-320 */
        match = op - offset;
        length = token & 15;

    copy_match:
        if (length == 15) {                                            /* :326-334 */
            for (;;) {
                uint32_t s = src[ip++];
                length += s;
                if (ip >= iend - K_LASTLITERALS + 1) goto output_error;
                if (s != 255) break;
            }
        }
        length += K_MINMATCH;                                          /* :336 */
        if (checkOffset && match + dictSize < 0) goto output_error;    /* :338 */

        if (extDict && match < 0) {                                    /* :341-378 */
            if (op + length > oend - K_LASTLITERALS) {
                if (partial) length = (oend - op) < length ? (oend - op) : length;
                else goto output_error;
            }
            for (int64_t i = 0; i < length; i++) {                     /* bytes of the window [dict | dst] */
                const int64_t j = match + i;
                dst[op + i] = j < 0 ? dict[dictSize + j] : dst[j];
            }
            op += length;
            continue;
        }
        if (match < 0) {
            /* no dictionary memory behind dst (checkOffset off can only happen with a dictionary):
             * the reference would read foreign memory; unreachable through LZ4Codec */
            goto output_error;
        }

        cpy = op + length;                                             /* :383 */
        if (partial && cpy > oend - 12) {                              /* :387-406 */
            const int64_t mlen = length < oend - op ? length : oend - op;
            for (int64_t i = 0; i < mlen; i++) dst[op + i] = offset ? dst[match + i] : 0;
            op += mlen;
            if (op == oend) break;
            continue;
        }
        if (cpy > oend - 12 && cpy > oend - K_LASTLITERALS) goto output_error;   /* :427-433 */
        if (offset == 0) memset(dst + op, 0, (size_t)length);
        else for (int64_t i = 0; i < length; i++) dst[op + i] = dst[match + i];
        op = cpy;                                                      /* :450 */
    }
    return (int)op;                                                    /* :454-457 */

output_error:
    return (int)(-ip) - 1;                                             /* :465 */
}

/* LZ4_decompress_safe_usingDict -- Engine/x64/LL64.dec.cs:523-546 */
int k4o_decompress_safe_usingDict(const uint8_t *src, int n, uint8_t *dst, int cap,
                                  const uint8_t *dict, int dictSize)
{
    if (dictSize == 0 || dict == NULL) return k4o_decompress_safe(src, n, dst, cap);
    return decompress_general(src, n, dst, cap, 0, dict, dictSize);
}

/* LZ4_decompress_safe_partial as called by LLxx.cs:29-39 (dstCapacity == targetOutputSize) */
int k4o_decompress_safe_partial(const uint8_t *src, int n, uint8_t *dst, int target)
{
    return decompress_general(src, n, dst, target, 1, NULL, 0);
}

/* LZ4Codec.Decode(byte*,int,byte*,int,byte*,int) -- LZ4Codec.cs:144-157 */
int k4o_codec_decode_dict(const uint8_t *src, int n, uint8_t *dst, int cap, const uint8_t *dict, int dictSize)
{
    if (n <= 0) return 0;
    int r = k4o_decompress_safe_usingDict(src, n, dst, cap, dict, dictSize);
    return r <= 0 ? -1 : r;
}

/* LZ4Codec.PartialDecode(byte*,int,byte*,int) -- LZ4Codec.cs:123-134 */
int k4o_codec_partial_decode(const uint8_t *src, int n, uint8_t *dst, int target)
{
    if (n <= 0) return 0;
    int r = k4o_decompress_safe_partial(src, n, dst, target);
    return r <= 0 ? -1 : r;
}

/* LZ4Codec.Encode(byte*,int,byte*,int,LZ4Level) -- LZ4Codec.cs:40-52.
 * level >= 3 (HC) is outside the hot path: -2 tells the caller to delegate. */
int k4o_codec_encode(const uint8_t *src, int n, uint8_t *dst, int cap, int level, int enforce32)
{
    if (n <= 0) return 0;
    if (level >= 3) return -2;
    int r = k4o_compress_fast(src, n, dst, cap, enforce32);
    return r <= 0 ? -1 : r;
}

/* LZ4Codec.Decode(byte*,int,byte*,int) -- LZ4Codec.cs:104-115 */
int k4o_codec_decode(const uint8_t *src, int n, uint8_t *dst, int cap)
{
    if (n <= 0) return 0;
    int r = k4o_decompress_safe(src, n, dst, cap);
    return r <= 0 ? -1 : r;
}

/* ---- LZ4Pickler, byte[] variant ---------------------------------------------------- */

/* LZ4Pickler.pickle.cs:224-225 (EffectiveSizeOf) */
static int size_of_diff(int v) { return (v > 0xffff || v < 0) ? 4 : (v > 0xff ? 2 : 1); }

/* Upper bound of a pickle: 1 + n (raw form) is the largest (compressed form is < n + 1). */
int k4o_pickle_bound(int n) { return n <= 0 ? 0 : n + 1; }

/*
 * LZ4Pickler.Pickle(ReadOnlySpan<byte>, level) -- LZ4Pickler.pickle.cs:51-106.
 * scratch capacity = 1024 if n <= 1024 else n (:57-67; PinnedMemory.cs:33-37,108).
 * `scratch` must hold max(n,1024) bytes.  Returns pickle length (0 for empty input).
 */
int k4o_pickle(const uint8_t *src, int n, uint8_t *dst, uint8_t *scratch, int level)
{
    if (n == 0) return 0;                                              /* :54 */
    int cap = n <= 1024 ? 1024 : n;
    int enc = k4o_codec_encode(src, n, scratch, cap, level, 0);        /* :83 */
    if (enc <= 0 || enc >= n) {                                        /* :85-94 */
        dst[0] = 0;
        memcpy(dst + 1, src, (size_t)n);
        return 1 + n;
    }
    int diff = n - enc;                                                /* :203-212 */
    int k = size_of_diff(diff);
    dst[0] = (uint8_t)(((k == 4 ? 3 : k) & 3) << 6);                   /* :221-228 */
    for (int i = 0; i < k; i++) dst[1 + i] = (uint8_t)((uint32_t)diff >> (8 * i));
    memcpy(dst + 1 + k, scratch, (size_t)enc);
    return 1 + k + enc;
}

/*
 * LZ4Pickler.Pickle<TBufferWriter>(source, writer, level) -- LZ4Pickler.pickle.cs:113-148.
 * Differs from the byte[] variant in two ways that change the bytes: the header width is fixed
 * BEFORE encoding from the full length (GetPessimisticHeaderSize, :129,161-165), and the payload
 * is encoded straight into the writer's span with capacity n (:130-133), not into the 1024-byte
 * minimum scratch.  `dst` must hold k4o_pickle_writer_bound(n) bytes.  Returns bytes advanced.
 */
int k4o_pickle_writer_bound(int n) { return n <= 0 ? 0 : 1 + size_of_diff(n) + n; }

int k4o_pickle_writer(const uint8_t *src, int n, uint8_t *dst, int level)
{
    if (n == 0) return 0;                                              /* :122 */
    const int headerSize = 1 + size_of_diff(n);                        /* :129 */
    int enc = k4o_codec_encode(src, n, dst + headerSize, n, level, 0); /* :132-133 */
    if (enc <= 0 || enc >= n) {                                        /* :135-140 */
        dst[0] = 0;
        memmove(dst + 1, src, (size_t)n);
        return 1 + n;
    }
    const int diff = n - enc, k = headerSize - 1;                      /* :203-212 with the pessimistic width */
    dst[0] = (uint8_t)(((k == 4 ? 3 : k) & 3) << 6);
    for (int i = 0; i < k; i++) dst[1 + i] = (uint8_t)((uint32_t)diff >> (8 * i));
    return headerSize + enc;                                           /* :146 */
}

/*
 * DecodeHeaderV0 + UnpickledSize -- LZ4Pickler.unpickle.cs:131-158,83-92.
 * Returns the unpickled size, or K4O_PICKLE_CORRUPT where the reference throws
 * InvalidDataException (bad version :131-135, truncated header :142-143,153).
 */
int k4o_unpickled_size(const uint8_t *src, int n)
{
    if (n == 0) return 0;                                              /* :41 */
    uint8_t h = src[0];
    if ((h & 7) != 0) return K4O_PICKLE_CORRUPT;
    int k = (h >> 6) & 3; if (k == 3) k = 4;
    int datalen = n - 1 - k;
    if (datalen < 0) return K4O_PICKLE_CORRUPT;
    uint32_t diff = 0;
    for (int i = 0; i < k; i++) diff |= (uint32_t)src[1 + i] << (8 * i);
    return datalen + (int)diff;
}

/*
 * LZ4Pickler.Unpickle(ReadOnlySpan<byte>, Span<byte>) -- LZ4Pickler.unpickle.cs:99-129.
 * dstLen must equal the expected size (:115-117).  Returns bytes produced, or
 * K4O_PICKLE_CORRUPT where the reference throws.
 */
int k4o_unpickle(const uint8_t *src, int n, uint8_t *dst, int dstLen)
{
    if (n == 0) return 0;
    int expected = k4o_unpickled_size(src, n);
    if (expected < 0) return K4O_PICKLE_CORRUPT;
    uint8_t h = src[0];
    int k = (h >> 6) & 3; if (k == 3) k = 4;
    uint32_t diff = 0;
    for (int i = 0; i < k; i++) diff |= (uint32_t)src[1 + i] << (8 * i);
    if (dstLen != expected) return K4O_PICKLE_CORRUPT;
    if (diff == 0) {                                                   /* :119-123 */
        memcpy(dst, src + 1 + k, (size_t)(n - 1 - k));
        return expected;
    }
    int dec = k4o_codec_decode(src + 1 + k, n - 1 - k, dst, dstLen);   /* :125 */
    if (dec != expected) return K4O_PICKLE_CORRUPT;                    /* :126-128 */
    return expected;
}

/* ---- XXH32 (LZ4 Frame checksums) --------------------------------------------------------------
 * Restates /root/reference/orig/lib/xxhash.c:263-390 (XXH32_round, XXH32_avalanche, XXH32_finalize,
 * XXH32_endian_align); the reference consumes it through NuGet K4os.Hash.xxHash
 * (Streams/Frames/LZ4FrameWriter.cs:162-181, LZ4FrameReader.cs:114-134). */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

uint32_t k4o_xxh32(const uint8_t *p, size_t len, uint32_t seed)
{
    const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    size_t i = 0;
    uint32_t h;
    if (len >= 16) {
        uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        for (; i + 16 <= len; i += 16) {
            v1 = rotl32(v1 + rd32(p + i) * P2, 13) * P1;
            v2 = rotl32(v2 + rd32(p + i + 4) * P2, 13) * P1;
            v3 = rotl32(v3 + rd32(p + i + 8) * P2, 13) * P1;
            v4 = rotl32(v4 + rd32(p + i + 12) * P2, 13) * P1;
        }
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else {
        h = seed + P5;
    }
    h += (uint32_t)len;
    for (; i + 4 <= len; i += 4) h = rotl32(h + rd32(p + i) * P3, 17) * P4;
    for (; i < len; i++) h = rotl32(h + p[i] * P5, 11) * P1;
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}
