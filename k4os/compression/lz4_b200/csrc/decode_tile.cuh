// decode_tile.cuh -- launch policy of the batched decoder.
// v0: every block goes through the warp-per-block global-memory decoder (decode_generic.cuh).
#pragma once
#include "common.cuh"
#include "decode_generic.cuh"

namespace k4 {

inline void decode_tile_set_attrs() {}

// returns the number of kernels launched
inline int decode_tile_launch(const uint8_t* srcBase, const int64_t* srcOff, const int32_t* srcLen,
                              uint8_t* dstBase, const int64_t* dstOff, const int32_t* dstCap,
                              int32_t* outLen, int n, cudaStream_t st) {
    const int ctas = (n + 3) / 4;
    decode_generic_kernel<<<ctas, 128, 0, st>>>(srcBase, srcOff, srcLen, dstBase, dstOff, dstCap,
                                                outLen, n, nullptr, n);
    return 1;
}

}  // namespace k4
