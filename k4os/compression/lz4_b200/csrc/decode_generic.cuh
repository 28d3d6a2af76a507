// decode_generic.cuh -- size-agnostic LZ4 block decoder: one warp per block, streams read and
// written in global memory.  It is the path for blocks that do not fit the shared-memory
// tile decoder (decode_tile.cuh) and the warp-level engine behind unpickle.
//
// Semantics = LZ4_decompress_safe as reached from LZ4Codec.Decode:
//   /root/reference/src/K4os.Compression.LZ4/Engine/x64/LL64.dec.cs:124-467 (generic loop,
//   endOnInputSize / full / noDict / lowPrefix = dst) and :469-477; every accept/reject test
//   is evaluated in the reference's order so that the returned value matches for malformed
//   input as well.  Control is warp-uniform: all lanes track (ip, op); the copies are
//   lane-parallel.  A match with offset 0 (accepted by the reference, content unspecified)
//   produces zero bytes.
#pragma once
#include "common.cuh"

namespace k4 {

// dst[op .. op+len) = src bytes; lanes stride by 32.
__device__ __forceinline__ void warp_copy_in(uint8_t* __restrict__ d, const uint8_t* __restrict__ s,
                                             int len, int lane) {
    for (int i = lane; i < len; i += 32) d[i] = __ldg(s + i);
}

// LZ77 match copy inside dst.  For an overlapping match (offset < len) the result is periodic
// with period `offset`, so every byte is fetched from the already-final window
// [match, match+offset): no intra-copy dependency between lanes.
__device__ __forceinline__ void warp_copy_match(uint8_t* dst, int64_t op, int64_t match,
                                                int len, int offset, int lane) {
    __syncwarp();   // earlier stores of other lanes must be visible
    if (offset == 0) {
        for (int i = lane; i < len; i += 32) dst[op + i] = 0;
    } else if (offset >= len) {
        for (int i = lane; i < len; i += 32) dst[op + i] = dst[match + i];
    } else {
        for (int i = lane; i < len; i += 32) dst[op + i] = dst[match + (i % offset)];
    }
}

// Returns bytes written (>= 0) or a negative value on malformed input / insufficient room.
__device__ int decode_block_warp(const uint8_t* __restrict__ src, int n,
                                 uint8_t* __restrict__ dst, int cap) {
    const int lane = lane_id();
    int64_t ip = 0, op = 0;
    const int64_t iend = n, oend = cap;
    const int64_t shortiend = iend - 16;      // LL64.dec.cs:152
    const int64_t shortoend = oend - 32;      // LL64.dec.cs:153

    if (cap == 0) return (n == 1 && __ldg(src) == 0) ? 0 : -1;   // :162-168
    if (n == 0) return -1;                                       // :172

    for (;;) {
        const uint32_t token = __ldg(src + ip); ip++;            // :177
        int64_t len = token >> 4;
        const bool shortcut = (len != 15) && (ip < shortiend) && (op <= shortoend);   // :191-193
        if (!shortcut) {
            if (len == 15) {                                     // :228-243, LL.tools.cs:165-193
                if (ip >= iend - 15) return -1;                  // initial_error
                for (;;) {
                    uint32_t s = __ldg(src + ip); ip++;
                    len += s;
                    if (ip >= iend - 15) break;                  // loop_error: not fatal here
                    if (s != 255) break;
                }
            }
            const int64_t cpy = op + len;                        // :246
            if (cpy > oend - MFLIMIT || ip + len > iend - (2 + 1 + LASTLITERALS)) {
                if (ip + len != iend || cpy > oend) return -1;   // :291-294
                warp_copy_in(dst + op, src + ip, (int)len, lane);
                return (int)(op + len);                          // :304-307, :454-457
            }
        }
        warp_copy_in(dst + op, src + ip, (int)len, lane);        // :196-200 / :311-314
        ip += len; op += len;

        const int offset = (int)__ldg(src + ip) | ((int)__ldg(src + ip + 1) << 8);   // :205 / :318
        ip += 2;
        const int64_t match = op - offset;
        len = token & 15;                                        // :204 / :323

        if (shortcut && len != 15 && offset >= 8 && match >= 0) {   // :211-220
            len += MINMATCH;
            warp_copy_match(dst, op, match, (int)len, offset, lane);
            op += len;
            continue;
        }
        if (len == 15) {                                         // :326-334
            for (;;) {
                uint32_t s = __ldg(src + ip); ip++;
                len += s;
                if (ip >= iend - LASTLITERALS + 1) return -1;    // any overrun is fatal
                if (s != 255) break;
            }
        }
        len += MINMATCH;                                         // :336
        if (match < 0) return -1;                                // :338
        const int64_t cpy = op + len;                            // :383
        if (cpy > oend - LASTLITERALS) return -1;                // :427-433
        warp_copy_match(dst, op, match, (int)len, offset, lane);
        op = cpy;                                                // :450
    }
}

// LZ4Codec.Decode post-processing (LZ4Codec.cs:104-115): len <= 0 -> 0 ; result <= 0 -> -1.
__device__ __forceinline__ int codec_decode_warp(const uint8_t* src, int n, uint8_t* dst, int cap) {
    if (n <= 0) return 0;
    int r = decode_block_warp(src, n, dst, cap);
    return r <= 0 ? -1 : r;
}

__global__ void __launch_bounds__(128)
decode_generic_kernel(const uint8_t* __restrict__ srcBase, const int64_t* __restrict__ srcOff,
                      const int32_t* __restrict__ srcLen, uint8_t* __restrict__ dstBase,
                      const int64_t* __restrict__ dstOff, const int32_t* __restrict__ dstCap,
                      int32_t* __restrict__ outLen, int nBlocks,
                      const int32_t* __restrict__ workList /* may be null: identity */, int nWork) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (w >= nWork) return;
    const int b = workList ? workList[w] : w;
    if (b < 0 || b >= nBlocks) return;
    int r = codec_decode_warp(srcBase + srcOff[b], srcLen[b], dstBase + dstOff[b], dstCap[b]);
    if (lane_id() == 0) outLen[b] = r;
}

}  // namespace k4
