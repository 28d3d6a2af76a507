// decode_parse.cuh -- pass 1 of the batched decoder: one THREAD per block walks the LZ4 token
// chain (the only inherently serial part of decoding) and validates it exactly as the reference
// does, emitting one 32-bit descriptor per sequence for the copy kernel.
//
// Why it looks the way it does (measured on B200, profiles/): a naive per-thread walk over global
// memory runs at ~1 us per sequence -- 32 lanes stream 32 different blocks, nearly every step one
// of them misses L1, and divergent branches serialise those latencies.  So
//   * every lane owns a 256-byte ring in shared memory that it keeps filled with 16-byte
//     cp.async copies issued ~8 steps ahead (one commit group per step, wait_group 8: data is
//     consumed only after it has had 8 steps to land; a lane whose bytes are not there yet simply
//     idles for that step instead of blocking the warp), and
//   * the walk is a branch-light state machine executed in lock step by all lanes:
//     TOKEN -> [LITVLE] -> LITEND -> OFFSET -> [MATCHVLE] -> MATCHCHK -> SEQEND -> TOKEN ...
//     a lane runs through as many states per step as its ring has bytes for (normally a whole
//     sequence).
//
// Semantics: LZ4_decompress_generic(endOnInputSize, full, noDict) --
//   /root/reference/src/K4os.Compression.LZ4/Engine/x64/LL64.dec.cs:124-467; every accept/reject
//   test is evaluated on the same values and in the same order, including the two-stage shortcut
//   (:191-225) whose eligibility depends on the output position, and LZ4_readVLE's early stop
//   (Engine/LL.tools.cs:165-193).  Results follow LZ4Codec.Decode (LZ4Codec.cs:104-115).
#pragma once
#include "common.cuh"

namespace k4 {

constexpr int TILE_BYTES = 65536;
constexpr int DESC_CAP = 16640;          // >= 65536/4 + 2 descriptors per block, multiple of 128
constexpr int MAX_BATCHES = (DESC_CAP + 31) / 32;   // 520
constexpr int ST_OK = 1, ST_DONE = 0, ST_FALLBACK = 2;   // BlockInfo.status

struct BlockInfo {
    int32_t nseq;             // descriptors emitted
    int32_t status;           // ST_OK: copy kernel materialises `outLen` bytes; ST_DONE: result already
                              // final (empty / error); ST_FALLBACK: generic decoder takes the block
    int32_t outLen;           // decoded size when ST_OK
    int32_t lastIsTerminal;   // the last descriptor is the final literal-only sequence
};

#ifndef K4_PARSE_BULK
#define K4_PARSE_BULK 1
#endif
constexpr int PARSE_THREADS = 128;
constexpr int PR_RING = 256;             // bytes of ring per lane
constexpr int PR_NCH = PR_RING / 16;     // 16-byte chunks per ring
constexpr int PR_G = 8;                  // commit groups (= steps) a copy is given to land
// Lane rings are 256 + 16 bytes apart: with a 256-byte stride all 32 lanes of a 16-byte cp.async hit
// the same four banks (32-way conflict, measured: half of all stall samples); 272 = 68 words spreads
// eight lanes over the 32 banks, the minimum for a 512-byte warp access.
constexpr int PR_STRIDE = PR_RING + 16;

enum ParseState : int { PS_TOKEN = 0, PS_LITVLE, PS_LITEND, PS_OFFSET, PS_MATCHVLE, PS_MATCHCHK, PS_SEQEND,
                        PS_FINAL_OK, PS_FINAL_ERR, PS_FINAL_FALLBACK };

__global__ void __launch_bounds__(PARSE_THREADS)
decode_parse_kernel(const uint8_t* __restrict__ srcBase, const int64_t* __restrict__ srcOff,
                    const int32_t* __restrict__ srcLen, const int32_t* __restrict__ dstCap,
                    int32_t* __restrict__ outLen, BlockInfo* __restrict__ info,
                    uint32_t* __restrict__ descs, int first, int count) {
    __shared__ __align__(16) uint8_t rings[PARSE_THREADS * PR_STRIDE];
#if K4_PARSE_BULK
    __shared__ unsigned long long bars[PARSE_THREADS * 2];   // one mbarrier per 128-byte ring half per lane
#endif
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < count;
    const int b = first + (live ? t : 0);
    int n = live ? srcLen[b] : 0;
    const int cap = live ? dstCap[b] : 0;
    BlockInfo bi; bi.nseq = 0; bi.status = ST_DONE; bi.outLen = 0; bi.lastIsTerminal = 0;

    int state = PS_TOKEN;
    if (!live) state = PS_FINAL_ERR;
    else if (n <= 0) { outLen[b] = 0; info[t] = bi; state = PS_FINAL_ERR; n = 0; }          // LZ4Codec.cs:108-109
    else if (cap <= 0) { outLen[b] = -1; info[t] = bi; state = PS_FINAL_ERR; }              // LL64.dec.cs:162-168
    const bool skipFinal = (state == PS_FINAL_ERR);      // result already written above

    const uint8_t* src = srcBase + (live ? srcOff[b] : 0);
    uint32_t* __restrict__ d = descs + (size_t)(live ? t : 0) * DESC_CAP;
    const int shift = (int)(reinterpret_cast<uintptr_t>(src) & 15);
    const uint8_t* gbase = src - shift;                       // 16-byte aligned
#if !K4_PARSE_BULK
    const int nChunks = (n + shift + 15) >> 4;
#endif
    uint8_t* ring = rings + threadIdx.x * PR_STRIDE;
    const uint32_t ringS = (uint32_t)__cvta_generic_to_shared(ring);

    int ip = 0, op = 0;
    const int iend = n, oend = cap;
    const int shortiend = iend - 16, shortoend = oend - 32;                                  // :152-153
#if !K4_PARSE_BULK
    int req = 0;                      // next 16-byte chunk to request (absolute index from gbase)
#endif
    int nseq = 0, result = -1;
    int tokPos = 0, seqOut = 0, len = 0, match = 0;
    uint32_t token = 0;
    bool shortcut = false;

#define PR_A(pos) ((pos) + shift)
#define PR_RD(pos) ((uint32_t)ring[PR_A(pos) & (PR_RING - 1)])

#if K4_PARSE_BULK
    // Ring = two 128-byte halves, each filled by ONE bulk copy (cp.async.bulk global -> shared: a whole
    // line per request instead of eight scattered 16-byte pieces) that completes on the lane's own
    // mbarrier; a lane polls its barriers without blocking and idles for a step when the bytes it
    // needs have not landed.  Invariant: at most two halves in flight, on alternating slots, and a
    // slot is only re-armed after its previous copy has been OBSERVED complete (hReq - hRdy < 2).
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bars[threadIdx.x * 2]);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0 + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int stagedEnd = (n + shift + 15) & ~15;             // bytes worth fetching, from gbase
    int hReq = 0;                     // next 128-byte half to request (absolute index from gbase)
    int hRdy = 0;                     // halves [.., hRdy) are known to have landed (hRdy <= hReq)
    uint32_t phase = 0;               // bit s: parity of the next completion to wait for on slot s
#endif
    while (__any_sync(FULL, state < PS_FINAL_OK)) {
#if K4_PARSE_BULK
        int rdyEnd;
        {
            const int curHalf = PR_A(ip) >> 7;
#pragma unroll
            for (int u = 0; u < 2; u++) {                           // landed?  oldest first
                if (hRdy < hReq) {
                    const int slot = hRdy & 1;
                    uint32_t okk;
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(okk) : "r"(bar0 + 8u * slot), "r"((phase >> slot) & 1u) : "memory");
                    if (okk) { hRdy++; phase ^= 1u << slot; }
                }
            }
            // jumped past everything requested (long literal run): once nothing is in flight, restart
            // at the half holding ip
            if (curHalf >= hReq && hRdy == hReq) hReq = hRdy = curHalf;
            if (state < PS_FINAL_OK && hReq >= curHalf && hReq <= curHalf + 1 && hReq - hRdy < 2 &&
                (hReq << 7) < stagedEnd) {
                int bytes = stagedEnd - (hReq << 7);
                if (bytes > 128) bytes = 128;
                const int slot = hReq & 1;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar0 + 8u * slot), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(ringS + 128u * slot), "l"(gbase + ((size_t)hReq << 7)), "r"(bytes), "r"(bar0 + 8u * slot) : "memory");
                hReq++;
            }
            rdyEnd = (hRdy << 7) < stagedEnd ? (hRdy << 7) : (stagedEnd + 4096);
            if (hRdy <= curHalf) rdyEnd = 0;                        // the half holding ip itself is not there yet
        }
#else
        // ---- keep the ring filled: at most one 16-byte copy per step per lane -------------------
        const int curChunk = PR_A(ip) >> 4;
        if (req < curChunk) req = curChunk;                   // jumped over a long literal run
        if (state < PS_FINAL_OK && req < curChunk + PR_NCH) {
            if (req < nChunks) {
                const uint32_t sdst = ringS + (uint32_t)((req & (PR_NCH - 1)) << 4);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(sdst), "l"(gbase + ((size_t)req << 4)) : "memory");
            }
            req++;                                            // past the end the requests are virtual
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" :: "n"(PR_G) : "memory");
        const int rdyEnd = (req - PR_G) << 4;                 // absolute byte bound of landed data
#endif
#define PR_READY(pos, k) (PR_A(pos) + (k) <= rdyEnd)

        // ---- fast path: the ordinary sequence (lengths with at most two extension bytes), nothing
        // exceptional.  It commits only when the reference would walk straight through -- no terminal
        // run, no error, not close to either buffer end where the extension-byte rules differ --
        // otherwise it leaves the lane untouched for the general machine below.
        bool needSlow = (state != PS_TOKEN && state != PS_OFFSET && state < PS_FINAL_OK);
        if (state == PS_TOKEN && PR_READY(ip, 3)) {
            const uint32_t tk = PR_RD(ip);
            int l = (int)(tk >> 4), ipL = ip + 1;
            bool ok = true;
            if (l == 15) {                                       // 15 .. 524 literals: one or two length bytes
                const uint32_t e1 = PR_RD(ip + 1), e2 = PR_RD(ip + 2);
                ok = ip + 3 < iend - 15;                         // far from the end: no early stop (LL.tools.cs:165-193)
                l += (int)e1; ipL++;
                if (e1 == 255) { ok = ok && e2 != 255; l += (int)e2; ipL++; }
            }
            const bool sc = (l < 15) && (ip + 1 < shortiend) && (op <= shortoend);              // :191-193
            ok = ok && (sc || !(op + l > oend - MFLIMIT || ipL + l > iend - (2 + 1 + LASTLITERALS)));
            if (ok) { tokPos = ip; token = tk; shortcut = sc; seqOut = op; ip = ipL + l; op += l; state = PS_OFFSET; }
            else needSlow = true;
        }
        if (state == PS_OFFSET && PR_READY(ip, 4)) {
            const int offset = (int)(PR_RD(ip) | (PR_RD(ip + 1) << 8));
            const int mt = op - offset;
            int ml = (int)(token & 15), ipM = ip + 2;
            bool ok = true;
            if (!(shortcut && ml != 15 && offset >= 8 && mt >= 0)) {                             // :211-220
                if (ml == 15) {                                  // 19 .. 528: one or two length bytes
                    const uint32_t e1 = PR_RD(ip + 2), e2 = PR_RD(ip + 3);
                    ok = ip + 4 < iend - LASTLITERALS + 1;       // far from the end: no overrun error
                    ml += (int)e1; ipM++;
                    if (e1 == 255) { ok = ok && e2 != 255; ml += (int)e2; ipM++; }
                }
                ok = ok && (mt >= 0) && !(op + ml + MINMATCH > oend - LASTLITERALS);             // :338, :427-433
            }
            const int opN = op + ml + MINMATCH;
            ok = ok && opN <= TILE_BYTES && tokPos <= 65535;
            if (ok) { d[nseq++] = (uint32_t)tokPos | ((uint32_t)seqOut << 16); ip = ipM; op = opN; state = PS_TOKEN; }
            else needSlow = true;
        }
        if (!__any_sync(FULL, needSlow)) continue;

        // ---- the general state machine (reference line numbers: LL64.dec.cs) --------------------------
        if (state == PS_TOKEN && PR_READY(ip, 1)) {                                          // :177
            tokPos = ip;
            token = PR_RD(ip); ip++;
            len = (int)(token >> 4);
            shortcut = (len != 15) && (ip < shortiend) && (op <= shortoend);                // :191-193
            if (len == 15) state = (ip >= iend - 15) ? PS_FINAL_ERR : PS_LITVLE;            // :231-232
            else state = PS_LITEND;
        }
        if (state == PS_LITVLE) {                                                            // LL.tools.cs:165-193
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (state == PS_LITVLE && PR_READY(ip, 1)) {
                    const uint32_t s = PR_RD(ip); ip++;
                    len += (int)s;
                    if (ip >= iend - 15 || s != 255) state = PS_LITEND;   // loop_error only stops the sum
                }
            }
        }
        if (state == PS_LITEND) {
            if (!shortcut) {
                const int cpy = op + len;                                                    // :246
                if (cpy > oend - MFLIMIT || ip + len > iend - (2 + 1 + LASTLITERALS)) {
                    if (ip + len != iend || cpy > oend) state = PS_FINAL_ERR;                // :291-294
                    else if (cpy > TILE_BYTES || tokPos > 65535 || (len > 0 && op > 65535)) state = PS_FINAL_FALLBACK;
                    else {
                        if (len > 0) { d[nseq++] = (uint32_t)tokPos | ((uint32_t)op << 16); bi.lastIsTerminal = 1; }
                        result = cpy;                                                        // :454-457
                        state = PS_FINAL_OK;
                    }
                }
            }
            if (state == PS_LITEND) { seqOut = op; ip += len; op += len; state = PS_OFFSET; }
        }
        if (state == PS_OFFSET && PR_READY(ip, 2)) {                                         // :205 / :318
            const int offset = (int)(PR_RD(ip) | (PR_RD(ip + 1) << 8));
            ip += 2;
            match = op - offset;
            len = (int)(token & 15);
            if (shortcut && len != 15 && offset >= 8 && match >= 0) state = PS_SEQEND;       // :211-220
            else state = (len == 15) ? PS_MATCHVLE : PS_MATCHCHK;
        }
        if (state == PS_MATCHVLE) {                                                          // :326-334
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (state == PS_MATCHVLE && PR_READY(ip, 1)) {
                    const uint32_t s = PR_RD(ip); ip++;
                    len += (int)s;
                    if (ip >= iend - LASTLITERALS + 1) state = PS_FINAL_ERR;
                    else if (s != 255) state = PS_MATCHCHK;
                }
            }
        }
        if (state == PS_MATCHCHK) {
            if (match < 0 || op + len + MINMATCH > oend - LASTLITERALS) state = PS_FINAL_ERR;   // :338, :427-433
            else state = PS_SEQEND;
        }
        if (state == PS_SEQEND) {
            op += len + MINMATCH;
            if (op > TILE_BYTES || tokPos > 65535 || seqOut > 65535) state = PS_FINAL_FALLBACK;
            else { d[nseq++] = (uint32_t)tokPos | ((uint32_t)seqOut << 16); state = PS_TOKEN; }
        }
    }
#if K4_PARSE_BULK
    while (hRdy < hReq) {                                     // drain the bulk copies still in flight
        const int slot = hRdy & 1;
        uint32_t okk;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(okk) : "r"(bar0 + 8u * slot), "r"((phase >> slot) & 1u) : "memory");
        if (okk) { hRdy++; phase ^= 1u << slot; }
    }
#else
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
#undef PR_A
#undef PR_RD
#undef PR_READY

    if (!live || skipFinal) return;
    if (state == PS_FINAL_FALLBACK) { bi.status = ST_FALLBACK; info[t] = bi; return; }       // outLen by the fallback
    if (state != PS_FINAL_OK || result <= 0) { outLen[b] = -1; info[t] = bi; return; }       // LZ4Codec.cs:114
    outLen[b] = result;
    bi.nseq = nseq; bi.status = ST_OK; bi.outLen = result;
    info[t] = bi;
}

}  // namespace k4
