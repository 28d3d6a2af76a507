// synth.cuh -- deterministic LZ-friendly synthetic block generator (bench / test workloads).
// Our own generator (not a restatement of any reference code): every block is produced
// independently from (seed, blockIndex) by a 64-bit xorshift-multiply stream as an
// alternation of
//   * literal runs over a geometrically skewed alphabet (p ~ 1/8, about 4.3 bits/byte, so that
//     short accidental repeats are as frequent as in text-like data: a 64 KiB block at ratio
//     0.5 carries ~3 000 LZ4 sequences, i.e. sequence statistics comparable to lz4's datagen), and
//   * back-references (offset <= 32 KiB, within the block) of length 4..19 (7/8) or 19..530 (1/8),
// chosen with probability matchPermille/1000.  Integer-only, so the host and device builds
// produce identical bytes.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace k4 {

struct SynthRng {
    uint64_t s;
    __host__ __device__ explicit SynthRng(uint64_t seed, uint64_t index) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (index + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        s = (z ^ (z >> 31)) | 1ull;
    }
    __host__ __device__ inline uint32_t next() {
        s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
        return (uint32_t)((s * 0x2545F4914F6CDD1Dull) >> 32);
    }
};

__host__ __device__ inline int synth_ctz(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

__host__ __device__ inline uint32_t synth_len(SynthRng& r) {
    const uint32_t v = r.next();
    return ((v >> 7) & 7) ? (v & 15) : 15 + ((v >> 12) & 511);
}

__host__ __device__ inline void synth_block(uint8_t* out, int size, uint32_t matchPermille,
                                            uint64_t seed, uint64_t blockIndex) {
    SynthRng r(seed, blockIndex);
    int pos = 0;
    while (pos < size) {
        const uint32_t sel = r.next();
        if (pos > 0 && ((sel >> 8) % 1000u) < matchPermille) {
            int len = 4 + (int)synth_len(r);
            const uint32_t window = pos < 32768 ? (uint32_t)pos : 32768u;
            const int off = 1 + (int)(r.next() % window);
            if (len > size - pos) len = size - pos;
            for (int i = 0; i < len; i++, pos++) out[pos] = out[pos - off];
        } else {
            int len = (int)synth_len(r);
            if (len > size - pos) len = size - pos;
            for (int i = 0; i < len; i++) {
                const uint32_t v = r.next();
                const uint32_t g = v & (v >> 11) & (v >> 22) & 0x3FFu;   // each bit set with p = 1/8
                uint32_t k = g ? (uint32_t)synth_ctz(g) : 10u + ((v >> 27) & 31u);
                out[pos++] = (uint8_t)(48 + k);
            }
        }
    }
}

__global__ void synth_kernel(uint8_t* base, long long nBlocks, int blockSize, uint32_t matchPermille,
                             uint64_t seed, long long firstBlock) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlocks) return;
    synth_block(base + b * (long long)blockSize, blockSize, matchPermille, seed, (uint64_t)(firstBlock + b));
}

}  // namespace k4
