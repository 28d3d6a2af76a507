// xxh32.cuh -- XXH32, the checksum of the LZ4 Frame container: batched on the
// GPU for per-block checksums (four lanes per block: the four accumulators of XXH32 are the only
// parallelism inside one stream), and a plain host function for the frame header and the serial
// content checksum.
//
// The reference computes these through K4os.Hash.xxHash: Streams/Frames/LZ4FrameWriter.cs:162-181
// (BlockChecksum / ContentChecksum), LZ4FrameReader.cs:114-134; algorithm: orig/lib/xxhash.c:263-390.
#pragma once
#include "common.cuh"

namespace k4 {

constexpr uint32_t XXP1 = 2654435761u, XXP2 = 2246822519u, XXP3 = 3266489917u, XXP4 = 668265263u, XXP5 = 374761393u;

__host__ __device__ __forceinline__ uint32_t xx_rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__host__ __device__ __forceinline__ uint32_t xx_round(uint32_t acc, uint32_t w) { return xx_rotl(acc + w * XXP2, 13) * XXP1; }

// tail (< 16 bytes at p) + avalanche; h already holds the merged accumulators (or seed + P5) + total length
__host__ __device__ __forceinline__ uint32_t xx_finish(uint32_t h, const uint8_t* t, size_t rest) {
    size_t i = 0;
    for (; i + 4 <= rest; i += 4) {
        const uint32_t w = (uint32_t)t[i] | ((uint32_t)t[i + 1] << 8) | ((uint32_t)t[i + 2] << 16) | ((uint32_t)t[i + 3] << 24);
        h = xx_rotl(h + w * XXP3, 17) * XXP4;
    }
    for (; i < rest; i++) h = xx_rotl(h + (uint32_t)t[i] * XXP5, 11) * XXP1;
    h ^= h >> 15; h *= XXP2; h ^= h >> 13; h *= XXP3; h ^= h >> 16;
    return h;
}

// host: one stream, serially
inline uint32_t xxh32_host(const uint8_t* p, size_t len, uint32_t seed) {
    size_t i = 0;
    uint32_t h;
    auto rd32 = [&](size_t k) { return (uint32_t)p[k] | ((uint32_t)p[k + 1] << 8) | ((uint32_t)p[k + 2] << 16) | ((uint32_t)p[k + 3] << 24); };
    if (len >= 16) {
        uint32_t v1 = seed + XXP1 + XXP2, v2 = seed + XXP2, v3 = seed, v4 = seed - XXP1;
        for (; i + 16 <= len; i += 16) {
            v1 = xx_round(v1, rd32(i)); v2 = xx_round(v2, rd32(i + 4)); v3 = xx_round(v3, rd32(i + 8)); v4 = xx_round(v4, rd32(i + 12));
        }
        h = xx_rotl(v1, 1) + xx_rotl(v2, 7) + xx_rotl(v3, 12) + xx_rotl(v4, 18);
    } else {
        h = seed + XXP5;
    }
    h += (uint32_t)len;
    return xx_finish(h, p + i, len - i);
}

// device: four consecutive lanes per block, lane g owns accumulator g (word g of every 16-byte stripe)
__global__ void __launch_bounds__(128)
xxh32_batch_kernel(const uint8_t* __restrict__ base, const int64_t* __restrict__ off,
                   const int32_t* __restrict__ len, uint32_t seed, uint32_t* __restrict__ out, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = t >> 2, g = t & 3;
    const bool live = b < n;
    const int L = live ? (len[b] > 0 ? len[b] : 0) : 0;
    const uint8_t* p = base + (live ? off[b] : 0);
    uint32_t v = g == 0 ? seed + XXP1 + XXP2 : g == 1 ? seed + XXP2 : g == 2 ? seed : seed - XXP1;
    const int stripes = L >> 4;
    for (int s = 0; s < stripes; s++) v = xx_round(v, ldg_u32u(p + 16 * s + 4 * g));
    const unsigned quad = 0xFu << (threadIdx.x & 28);
    const uint32_t r = xx_rotl(v, g == 0 ? 1 : g == 1 ? 7 : g == 2 ? 12 : 18);
    uint32_t h = r + __shfl_xor_sync(quad, r, 1);
    h += __shfl_xor_sync(quad, h, 2);
    if (!live || g != 0) return;
    if (L < 16) h = seed + XXP5;
    h += (uint32_t)L;
    out[b] = xx_finish(h, p + 16 * stripes, (size_t)(L & 15));
}

}  // namespace k4
