// decode_tile.cuh -- pass 2 of the batched LZ4 block decoder for B200, and the launch policy.
//
//   pass 1  decode_parse.cuh   thread per block: validated sequence descriptors
//   pass 2  decode_copy_kernel one CTA (384 threads) per block.  The compressed block is pulled
//           into shared memory with TMA bulk copies (cp.async.bulk global -> shared, 8 KiB pieces,
//           one mbarrier each) and the 64 KiB output tile is built in shared memory as well, so
//           every byte the copy loop touches is a shared-memory access.  Warps take batches of
//           32 consecutive sequences, ONE SEQUENCE PER LANE: lane-parallel literal copies, then
//           lane-parallel match copies under EXACT BYTE-LEVEL dependencies: a bitmap with one bit
//           per output byte (8 KiB) records which bytes are final; a writer ORs its byte mask in
//           after a block-scope fence, a match is copied as soon as the bits of the bytes it reads
//           are set.  (Measured alternatives: per-batch done flags 38 ms, per-sequence flags through
//           a granule map 21 ms, this bitmap 13 ms per 4 GiB.)  Long runs are copied by the whole
//           warp.  The finished tile leaves through one TMA bulk store (cp.async.bulk shared -> global).
//
// Two instantiations: STAGE = 40 KiB (two CTAs per SM: 64 + 40 + 8 KiB; blocks whose compressed
// size fits) and STAGE = 66 KiB (one CTA per SM; everything else that still fits the 64 KiB tile).
// Blocks that do not fit the tile at all (decoded size > 64 KiB) go to the warp-per-block generic
// decoder inside the same launch.
//
// Reference semantics: /root/reference/src/K4os.Compression.LZ4/Engine/x64/LL64.dec.cs:124-477,
// Engine/LL.tools.cs:165-193 (LZ4_readVLE), LZ4Codec.cs:104-115.
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "decode_generic.cuh"
#include "decode_parse.cuh"

namespace k4 {

constexpr int SUB_BATCH = 65536;         // blocks per parse/copy kernel pair
#ifndef K4_COPY_THREADS
#define K4_COPY_THREADS 384
#endif
constexpr int COPY_THREADS = K4_COPY_THREADS;
constexpr int COPY_WARPS = COPY_THREADS / 32;
constexpr int STAGE_PIECE = 8192;        // bytes per TMA bulk load / mbarrier
constexpr int STAGE_SMALL = 40 * 1024;   // compressed bytes (incl. alignment slack) staged, 2 CTAs/SM
constexpr int STAGE_BIG = 66 * 1024;     // ... 1 CTA/SM; covers LZ4_compressBound(65536) + slack
constexpr int MAX_PIECES = (STAGE_BIG + STAGE_PIECE - 1) / STAGE_PIECE;   // 9

template <int STAGE>
struct CopySmem {
    uint8_t tile[TILE_BYTES];
    uint8_t stage[STAGE];
    volatile uint32_t ready[TILE_BYTES / 32];  // one bit per output byte: the byte is final in the tile
    unsigned long long bar[MAX_PIECES];        // one mbarrier per staged piece
};

__device__ __forceinline__ void tma_store_tile(uint8_t* gdst, const uint8_t* stile, int bytes) {
    // generic-proxy writes -> async proxy, then one bulk copy shared::cta -> global
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(stile);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(saddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}

// WAITMODE: 1 = poll with nanosleep back-off (product); 2 = no waiting (timing experiments only: wrong output)
template <int STAGE, int WAITMODE>
__global__ void __launch_bounds__(COPY_THREADS, (COPY_THREADS <= 512 ? 2 : 1))
decode_copy_kernel(const uint8_t* __restrict__ srcBase, const int64_t* __restrict__ srcOff,
                   const int32_t* __restrict__ srcLen, uint8_t* __restrict__ dstBase,
                   const int64_t* __restrict__ dstOff, const int32_t* __restrict__ dstCap,
                   int32_t* __restrict__ outLen, const BlockInfo* __restrict__ info,
                   const uint32_t* __restrict__ descs, int first) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    CopySmem<STAGE>& S = *reinterpret_cast<CopySmem<STAGE>*>(smem_raw);
    const int t = blockIdx.x;
    const int b = first + t;
    const BlockInfo bi = info[t];
    const int lane = threadIdx.x & 31;
    // the hardware arbiter favours high warp ids: give them the EARLIEST batches, the ones every
    // other warp may be waiting for
    const int warp = COPY_WARPS - 1 - (int)(threadIdx.x >> 5);

    if (bi.status == ST_DONE) return;
    const uint8_t* __restrict__ src = srcBase + srcOff[b];
    const int n = srcLen[b];
    const int shift = (int)(reinterpret_cast<uintptr_t>(src) & 15);
    const int staged = (n + shift + 15) & ~15;           // bytes pulled in, from the aligned base
    const bool big = staged > STAGE_SMALL;
    if (bi.status == ST_FALLBACK || staged > STAGE_BIG) {
        if (STAGE != STAGE_SMALL) return;                // handled once, by the small-stage launch
        if (warp == 0) {
            const int r = codec_decode_warp(src, n, dstBase + dstOff[b], dstCap[b]);
            if (lane == 0) outLen[b] = r;
        }
        return;
    }
    if (big != (STAGE == STAGE_BIG)) return;             // the other instantiation owns this block

    const uint32_t* __restrict__ d = descs + (size_t)t * DESC_CAP;
    const int nseq = bi.nseq;
    const int nbatch = (nseq + 31) >> 5;
    uint8_t* tile = S.tile;
    const uint8_t* stg = S.stage + shift;                // stg[p] == src[p]
    const int npieces = (staged + STAGE_PIECE - 1) / STAGE_PIECE;

    // ---- prologue: TMA loads of the compressed block, done flags, granule -> batch map ---------
    if (threadIdx.x == 0) {
        for (int i = 0; i < npieces; i++) {
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(&S.bar[i]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(a) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint8_t* gsrc = src - shift;
        for (int i = 0; i < npieces; i++) {
            const int off = i * STAGE_PIECE;
            const int bytes = (staged - off) < STAGE_PIECE ? (staged - off) : STAGE_PIECE;
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(&S.bar[i]);
            const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(S.stage + off);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(sdst), "l"(gsrc + off), "r"(bytes), "r"(a) : "memory");
        }
    }
    for (int i = threadIdx.x; i < TILE_BYTES / 32; i += COPY_THREADS) S.ready[i] = 0;
    __syncthreads();

    // ---- batches: one sequence per lane -----------------------------------------------------------
    int piecesSeen = 0;                                   // staged pieces this warp has waited for
    uint32_t descNext = (warp < nbatch && warp * 32 + lane < nseq) ? __ldg(d + warp * 32 + lane) : 0u;
    for (int bt = warp; bt < nbatch; bt += COPY_WARPS) {
        const int k = bt * 32 + lane;
        const bool valid = k < nseq;
        const uint32_t desc = descNext;
        {   // prefetch this warp's next descriptors
            const int kn = k + COPY_WARPS * 32;
            descNext = (kn < nseq) ? __ldg(d + kn) : 0u;
        }
        const uint32_t tokPos = desc & 0xFFFFu;
        const int dst = (int)(desc >> 16);
        // the staged pieces holding this batch's sequence headers (a header = token + length bytes,
        // at most ~260 bytes for a 64 KiB block); literal extents are waited for below
        {
            const int lastTok = __reduce_max_sync(FULL, valid ? (int)tokPos : 0);
            int needPieces = (lastTok + 600 + shift + STAGE_PIECE - 1) / STAGE_PIECE;
            if (needPieces > npieces) needPieces = npieces;
            while (piecesSeen < needPieces) { mbar_wait(&S.bar[piecesSeen], 0); piecesSeen++; }
        }

        // sequence header (lengths already validated by the parse kernel)
        int lit = 0, ml = 0, offset = 0;
        uint32_t litSrc = 0;
        bool hasMatch = false;
        if (valid) {
            const uint32_t token = stg[tokPos];
            uint32_t p = tokPos + 1;
            lit = (int)(token >> 4);
            if (lit == 15) {                                   // LZ4_readVLE incl. its early stop
                for (;;) {
                    const uint32_t s = stg[p]; p++;
                    lit += (int)s;
                    if ((int)p >= n - 15) break;
                    if (s != 255) break;
                }
            }
            litSrc = p;
            hasMatch = !(bi.lastIsTerminal && k == nseq - 1);
        }
        {   // every staged byte this batch reads (literals, then offset + match length bytes)
            const int endMax = __reduce_max_sync(FULL, valid ? (int)litSrc + lit + 300 : 0);
            int needPieces = (endMax + shift + STAGE_PIECE - 1) / STAGE_PIECE;
            if (needPieces > npieces) needPieces = npieces;
            while (piecesSeen < needPieces) { mbar_wait(&S.bar[piecesSeen], 0); piecesSeen++; }
        }
        if (hasMatch) {
            uint32_t p = litSrc + (uint32_t)lit;
            offset = (int)stg[p] | ((int)stg[p + 1] << 8);
            p += 2;
            ml = (int)(stg[tokPos] & 15);
            if (ml == 15) {
                for (;;) { const uint32_t s = stg[p]; p++; ml += (int)s; if (s != 255) break; }
            }
            ml += MINMATCH;
        }
        // helpers on the byte-readiness bitmap -----------------------------------------------------
        // mark [a, a+len) final (len <= 32: at most two words)
        auto publish_short = [&](const int a, const int len) {
            if (len > 0) {
                const unsigned long long m = ((len >= 64 ? 0ull : (1ull << len)) - 1ull) << (a & 31);
                uint32_t* w = const_cast<uint32_t*>(&S.ready[a >> 5]);
                atomicOr(w, (uint32_t)m);
                if ((uint32_t)(m >> 32)) atomicOr(w + 1, (uint32_t)(m >> 32));
            }
        };
        // are all bytes of [a, a+len) final? (len <= 32)
        auto ready_short = [&](const int a, const int len) -> bool {
            if (len <= 0) return true;
            const unsigned long long m = ((1ull << len) - 1ull) << (a & 31);
            const uint32_t lo32 = (uint32_t)m, hi32 = (uint32_t)(m >> 32);
            bool r = (S.ready[a >> 5] & lo32) == lo32;
            if (hi32) r = r && ((S.ready[(a >> 5) + 1] & hi32) == hi32);
            return r;
        };
        // whole-warp versions for runs of any length (uniform arguments)
        auto publish_long = [&](const int a, const int len) {
            const int w0 = a >> 5, w1 = (a + len - 1) >> 5;
            for (int w = w0 + lane; w <= w1; w += 32) {
                const int b0 = (w << 5) > a ? (w << 5) : a;
                const int b1 = ((w + 1) << 5) < (a + len) ? ((w + 1) << 5) : (a + len);
                const uint32_t m = (uint32_t)(((1ull << (b1 - b0)) - 1ull) << (b0 & 31));
                atomicOr(const_cast<uint32_t*>(&S.ready[w]), m);
            }
        };
        auto ready_long = [&](const int a, const int len) -> bool {
            const int w0 = a >> 5, w1 = (a + len - 1) >> 5;
            bool r = true;
            for (int w = w0 + lane; w <= w1; w += 32) {
                const int b0 = (w << 5) > a ? (w << 5) : a;
                const int b1 = ((w + 1) << 5) < (a + len) ? ((w + 1) << 5) : (a + len);
                const uint32_t m = (uint32_t)(((1ull << (b1 - b0)) - 1ull) << (b0 & 31));
                r = r && ((S.ready[w] & m) == m);
            }
            return __all_sync(FULL, r);
        };

        // ---- literals: short runs lane-parallel (all loads, then all stores), long runs by the warp -----
        {
            const int shortLit = lit < 15 ? lit : 0;
            const int mx = __reduce_max_sync(FULL, shortLit);
            if (mx > 0) {
                uint8_t v[14];
                if (mx <= 7) {
#pragma unroll
                    for (int j = 0; j < 7; j++) if (j < shortLit) v[j] = stg[litSrc + j];
#pragma unroll
                    for (int j = 0; j < 7; j++) if (j < shortLit) tile[dst + j] = v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 14; j++) if (j < shortLit) v[j] = stg[litSrc + j];
#pragma unroll
                    for (int j = 0; j < 14; j++) if (j < shortLit) tile[dst + j] = v[j];
                }
                __threadfence_block();
                publish_short(dst, shortLit);
            }
            unsigned longMask = __ballot_sync(FULL, lit >= 15);
            while (longMask) {
                const int l = __ffs(longMask) - 1;
                longMask &= longMask - 1;
                const uint32_t s0 = __shfl_sync(FULL, litSrc, l);
                const int d0 = __shfl_sync(FULL, dst, l);
                const int len = __shfl_sync(FULL, lit, l);
                for (int i = lane; i < len; i += 32) tile[d0 + i] = stg[s0 + i];
                __threadfence_block();
                __syncwarp();
                publish_long(d0, len);
            }
        }

        // ---- matches: exact byte-level dependencies ---------------------------------------------------
        const int mdst = dst + lit;
        const int msrc = mdst - offset;
        // bytes actually read: an overlapping match (offset < ml) only reads [msrc, mdst)
        const int srcLen = hasMatch ? (offset == 0 ? 0 : (offset < ml ? offset : ml)) : 0;
        unsigned backoff = 16u;
        bool pendS = hasMatch && ml <= 18;                 // short: lane-parallel
        unsigned pendL = __ballot_sync(FULL, hasMatch && ml > 18);   // long: whole warp, one at a time
        for (;;) {
            bool progress = false;
            // short matches whose source bytes are all final
            const bool go = pendS && (WAITMODE == 2 || ready_short(msrc, srcLen));
            if (__any_sync(FULL, go)) {
                __threadfence_block();                     // acquire the flagged bytes
                {   // non-overlapping: all loads first, then all stores
                    const int m1 = (go && offset >= ml) ? ml : 0;
                    const int mx = __reduce_max_sync(FULL, m1);
                    if (mx > 0) {
                        uint8_t v[18];
                        if (mx <= 8) {
#pragma unroll
                            for (int j = 0; j < 8; j++) if (j < m1) v[j] = tile[msrc + j];
#pragma unroll
                            for (int j = 0; j < 8; j++) if (j < m1) tile[mdst + j] = v[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < 18; j++) if (j < m1) v[j] = tile[msrc + j];
#pragma unroll
                            for (int j = 0; j < 18; j++) if (j < m1) tile[mdst + j] = v[j];
                        }
                    }
                }
                {   // overlapping (offset < ml) or offset 0: in-order byte loop per lane
                    const int m2 = (go && offset < ml) ? ml : 0;
                    const int mx = __reduce_max_sync(FULL, m2);
                    if (offset > 0) { for (int j = 0; j < mx; j++) if (j < m2) tile[mdst + j] = tile[msrc + j]; }
                    else            { for (int j = 0; j < mx; j++) if (j < m2) tile[mdst + j] = 0; }
                }
                __threadfence_block();
                if (go) publish_short(mdst, ml);
                pendS = pendS && !go;
                progress = true;
            }
            // long matches
            unsigned m = pendL;
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const int s0 = __shfl_sync(FULL, msrc, l);
                const int d0 = __shfl_sync(FULL, mdst, l);
                const int len = __shfl_sync(FULL, ml, l);
                const int sl = __shfl_sync(FULL, srcLen, l);
                if (!(WAITMODE == 2 || sl == 0 || ready_long(s0, sl))) continue;
                __threadfence_block();
                const int off = d0 - s0;
                if (off == 0)          for (int i = lane; i < len; i += 32) tile[d0 + i] = 0;
                else if (off >= len)   for (int i = lane; i < len; i += 32) tile[d0 + i] = tile[s0 + i];
                else                   for (int i = lane; i < len; i += 32) tile[d0 + i] = tile[s0 + (i % off)];
                __threadfence_block();
                __syncwarp();
                publish_long(d0, len);
                pendL &= ~(1u << l);
                progress = true;
            }
            if (!__any_sync(FULL, pendS) && !pendL) break;
            if (progress) backoff = 16u;
            else if (WAITMODE == 1) { __nanosleep(backoff); if (backoff < 512u) backoff <<= 1; }   // nothing ready: back off
        }
    }
    __syncthreads();

    // ---- tile -> global ------------------------------------------------------------------------------
    uint8_t* gdst = dstBase + dstOff[b];
    const int total = bi.outLen;
    if ((reinterpret_cast<uintptr_t>(gdst) & 15) == 0) {
        const int bulk = total & ~15;
        if (threadIdx.x == 0 && bulk > 0) tma_store_tile(gdst, tile, bulk);
        for (int i = bulk + threadIdx.x; i < total; i += COPY_THREADS) gdst[i] = tile[i];
    } else {
        // unaligned destination: byte head up to 4-byte alignment, then words built from the tile
        const int head = (int)((4 - (reinterpret_cast<uintptr_t>(gdst) & 3)) & 3);
        const int h = head < total ? head : total;
        for (int i = threadIdx.x; i < h; i += COPY_THREADS) gdst[i] = tile[i];
        const int words = (total - h) >> 2;
        uint32_t* g4 = reinterpret_cast<uint32_t*>(gdst + h);
        for (int i = threadIdx.x; i < words; i += COPY_THREADS) {
            const uint8_t* s = tile + h + 4 * i;
            g4[i] = (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24);
        }
        for (int i = h + 4 * words + threadIdx.x; i < total; i += COPY_THREADS) gdst[i] = tile[i];
    }
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
template <int STAGE, int W>
inline void decode_copy_launch_t(const uint8_t* srcBase, const int64_t* srcOff, const int32_t* srcLen,
                                 uint8_t* dstBase, const int64_t* dstOff, const int32_t* dstCap,
                                 int32_t* outLen, const BlockInfo* info, const uint32_t* descs,
                                 int first, int count, cudaStream_t st) {
    static bool attr[64] = {false};                       // function attributes are per device
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attr[dev]) {
        cudaFuncSetAttribute(decode_copy_kernel<STAGE, W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(CopySmem<STAGE>));
        attr[dev] = true;
    }
    decode_copy_kernel<STAGE, W><<<count, COPY_THREADS, sizeof(CopySmem<STAGE>), st>>>(
        srcBase, srcOff, srcLen, dstBase, dstOff, dstCap, outLen, info, descs, first);
}

inline void decode_tile_set_attrs() {}

// returns the number of kernels launched, or -1 on a CUDA error (cudaGetLastError has it)
inline int decode_tile_launch(const uint8_t* srcBase, const int64_t* srcOff, const int32_t* srcLen,
                              uint8_t* dstBase, const int64_t* dstOff, const int32_t* dstCap,
                              int32_t* outLen, int n, cudaStream_t st) {
    static const int variant = [] { const char* e = getenv("K4LZ4_COPY_VARIANT"); return e ? atoi(e) : 0; }();
    static const int subEnv = [] { const char* e = getenv("K4LZ4_SUB_BATCH"); return e ? atoi(e) : SUB_BATCH; }();
    const int sub = n < subEnv ? n : subEnv;
    {   // keep the stream-ordered pool's memory across calls (default: released at every sync)
        static bool poolSet[64] = {false};
        int dev = 0; cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !poolSet[dev]) {
            cudaMemPool_t pool; 
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long thr = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
            poolSet[dev] = true;
        }
    }
    uint32_t* descs = nullptr;
    BlockInfo* info = nullptr;
    if (cudaMallocAsync(&descs, (size_t)sub * DESC_CAP * sizeof(uint32_t), st) != cudaSuccess) return -1;
    if (cudaMallocAsync(&info, (size_t)sub * sizeof(BlockInfo), st) != cudaSuccess) { cudaFreeAsync(descs, st); return -1; }
    int launches = 0;
    for (int first = 0; first < n; first += sub) {
        const int count = (n - first) < sub ? (n - first) : sub;
        decode_parse_kernel<<<(count + PARSE_THREADS - 1) / PARSE_THREADS, PARSE_THREADS, 0, st>>>(
            srcBase, srcOff, srcLen, dstCap, outLen, info, descs, first, count);
        launches++;
#define K4_COPY(STG, W) decode_copy_launch_t<STG, W>(srcBase, srcOff, srcLen, dstBase, dstOff, dstCap, outLen, info, descs, first, count, st)
        switch (variant) {
        case 2: K4_COPY(STAGE_SMALL, 2); K4_COPY(STAGE_BIG, 2); launches += 2; break;   // no waiting (tools/dbench.py only)
        case 9: break;                                                                   // parse only (tools/dbench.py only)
        default: K4_COPY(STAGE_SMALL, 1); K4_COPY(STAGE_BIG, 1); launches += 2; break;
        }
#undef K4_COPY
    }
    cudaFreeAsync(descs, st);
    cudaFreeAsync(info, st);
    return launches;
}

}  // namespace k4
