// parse_table.cuh -- the jump table of the tile decoder's speculative parse.
//
// The decoder finds the token chain of a block by walking it from many arbitrary start bytes at
// once (decode_tile.cuh).  A hop of such a walk needs, at a position p taken as a token, only the
// distance to the next token.  For all but a few exotic sequences that distance depends on the
// token byte, the byte behind it and ONE further byte (the first match-length extension), so it is
// computed ONCE for every byte position of the stream, four positions per thread from two aligned
// words, and stored as one byte per position:
//
//     J[p] = next(p) - p     when the sequence at p is "plain":  literal-length extension (if any)
//                            of one byte < 224, match-length extension (if any) of one byte < 255,
//                            and the sequence ends before the end of the stream (so it is neither
//                            the terminal sequence nor malformed)
//     J[p] = 255             otherwise: the walker decodes p with the exact header code (seq_next)
//
// A hop then is one byte load and an add instead of ~45 instructions with two dependent loads.
// The table is exact where it is not 255: tests/test_parse_table.py compares every position of
// valid, mutated and random streams with a restatement of the reference's length decoding
// (LL64.dec.cs:191-246,300-336; LL.tools.cs:165-193).
//
// Plain C++ (no CUDA constructs) so that the same code compiles for the host-side test.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define K4_HD __host__ __device__ __forceinline__
#else
#define K4_HD inline
#endif

namespace k4 {

constexpr uint32_t JT_ESC = 255u;        // "decode this position with the exact header code"
constexpr int JT_MAXJ = 243;             // largest distance the table stores: 3 + 15 + 1 + 1 + 223

// Four table bytes for the stream positions p0 .. p0+3 whose bytes are w0 (little endian); w1 holds the
// four bytes behind them.  ld8(q) returns stream byte q for 0 <= q <= n (bytes at and behind n may be
// anything: every entry they influence becomes 255).  Positions outside [0, n) yield junk nobody reads.
template <class LD8>
K4_HD uint32_t jt_word(const uint32_t w0, const uint32_t w1, const int p0, const int n, LD8 ld8) {
    const uint32_t L4 = (w0 >> 4) & 0x0F0F0F0Fu, M4 = w0 & 0x0F0F0F0Fu;
    const uint32_t lf = ((L4 + 0x01010101u) >> 4) & 0x01010101u;          // 1 in every byte whose literal nibble is 15
    const uint32_t mf = ((M4 + 0x01010101u) >> 4) & 0x01010101u;          // ... whose match nibble is 15
    uint32_t e = ((w0 >> 8) | (w1 << 24)) & ((lf << 8) - lf);             // first literal extension byte where it counts
    const uint32_t big = (e & (e << 1) & (e << 2) & 0x80808080u) >> 7;    // 1 where that byte is >= 224 (255 continues)
    e &= ~((big << 8) - big);
    const uint32_t j4 = 0x03030303u + L4 + lf + mf + e;                    // token + lengths + offset; no byte exceeds 243
    uint32_t esc = big;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int b = 0; b < 4; b++) {
        if ((mf >> (8 * b)) & 1u) {                                        // the match extension is the sequence's last byte
            const int q = p0 + b + (int)((j4 >> (8 * b)) & 0xFFu) - 1;
            if (ld8(q < n ? q : n) == 255u) esc |= 1u << (8 * b);
        }
    }
    if (p0 + 3 + JT_MAXJ >= n) {                                           // near the end: a plain sequence ends BEFORE n
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int b = 0; b < 4; b++)
            if (p0 + b + (int)((j4 >> (8 * b)) & 0xFFu) >= n) esc |= 1u << (8 * b);
    }
    return j4 | ((esc << 8) - esc);
}

// Output bytes of the plain sequence at p (J[p] = j != 255): literals + match.
template <class LD8>
K4_HD uint32_t jt_outbytes(const int p, const uint32_t j, LD8 ld8) {
    const uint32_t tok = ld8(p);
    const uint32_t L = tok >> 4, M = tok & 15u;
    uint32_t o = L + M + 4u;
    if (L == 15u) o += ld8(p + 1);
    if (M == 15u) o += ld8(p + (int)j - 1);
    return o;
}

}  // namespace k4
