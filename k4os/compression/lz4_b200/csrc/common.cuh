// common.cuh -- shared device helpers and LZ4 block-format constants for libk4lz4.
// Constants restate Engine/LL.types.cs:50-78 of the reference (values only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace k4 {

constexpr int MINMATCH = 4;
constexpr int LASTLITERALS = 5;
constexpr int MFLIMIT = 12;
constexpr int MINLENGTH = 13;              // LZ4_minLength
constexpr int LIMIT_64K = 65536 + 11;      // LZ4_64Klimit: below it the u16 table is used
constexpr int MAX_DISTANCE = 65535;
constexpr int SKIP_TRIGGER = 6;
constexpr int MAX_INPUT_SIZE = 0x7E000000;
constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ inline int max_output_size(int n) {
    return n > MAX_INPUT_SIZE ? 0 : n + n / 255 + 16;
}

// Unaligned little-endian 32-bit read from global memory built from aligned words, so that
// no access ever touches a word that holds no byte of [p, p+4).
__device__ __forceinline__ uint32_t ldg_u32u(const uint8_t* p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t lo = __ldg(w);
    uint32_t hi = sh ? __ldg(w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}
__device__ __forceinline__ uint64_t ldg_u64u(const uint8_t* p) {
    return (uint64_t)ldg_u32u(p) | ((uint64_t)ldg_u32u(p + 4) << 32);
}

__device__ __forceinline__ uint32_t hash4(uint32_t v, int log) { return (v * 2654435761u) >> (32 - log); }
__device__ __forceinline__ uint32_t hash5(uint64_t v, int log) {
    return (uint32_t)(((v << 24) * 889523592379ull) >> (64 - log));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

}  // namespace k4
