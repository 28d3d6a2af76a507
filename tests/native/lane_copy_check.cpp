// Host-side check of csrc/lane_copy.cuh (word-granular per-lane copies of the tile decoder): 32 "lanes"
// copy runs of 0..32 bytes at random alignments inside one byte array; destinations of different lanes
// touch at byte boundaries, sources lie anywhere outside every destination.  The result must equal
// byte-wise copies, and no byte outside the destinations may change.  Driven by tests/test_parse_table.py.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <random>
#include <vector>
#include "../../k4os/compression/lz4_b200/csrc/lane_copy.cuh"

namespace {
struct Mem {
    std::vector<uint8_t>& b;
    long oob = 0;
    uint32_t ld8(uint32_t a) { if (a >= b.size()) { oob++; return 0; } return b[a]; }
    uint32_t ld32(uint32_t a) { if ((a & 3u) || a + 4 > b.size()) { oob++; return 0; } uint32_t v; memcpy(&v, &b[a], 4); return v; }
    void st8(uint32_t a, uint32_t v) { if (a >= b.size()) { oob++; return; } b[a] = (uint8_t)v; }
    void st32(uint32_t a, uint32_t v) { if ((a & 3u) || a + 4 > b.size()) { oob++; return; } memcpy(&b[a], &v, 4); }
};
}  // namespace

extern "C" long lc_check(int rounds, unsigned seed) {
    std::mt19937 rng(seed);
    long fails = 0;
    for (int r = 0; r < rounds; r++) {
        const int N = 8192;
        std::vector<uint8_t> buf(N), want;
        for (auto& x : buf) x = (uint8_t)rng();
        // destinations: 32 consecutive runs (touching, or with small gaps) in the upper half
        uint32_t d[32], s[32]; int len[32];
        uint32_t pos = 4096 + rng() % 64;
        for (int l = 0; l < 32; l++) {
            len[l] = (rng() % 5 == 0) ? 0 : (int)(rng() % (k4::LC_MAX + 1));
            if (rng() % 3 == 0) pos += rng() % 5;
            d[l] = pos; pos += (uint32_t)len[l];
            s[l] = 8 + rng() % (4096 - 8 - k4::LC_MAX);             // sources in the lower half: final
            if (rng() % 8 == 0 && len[l] > 0 && d[l] >= 4096 + 64) s[l] = d[0] - (uint32_t)len[l] - rng() % 3;   // ends right below the first destination
        }
        want = buf;
        for (int l = 0; l < 32; l++) for (int i = 0; i < len[l]; i++) want[d[l] + i] = buf[s[l] + i];
        int nwTop = 0;
        for (int l = 0; l < 32; l++) nwTop = std::max(nwTop, k4::lc_words(d[l], len[l]));
        Mem m{buf};
        for (int l = 0; l < 32; l++) k4::lc_copy(m, d[l], s[l], len[l], nwTop);
        if (m.oob || buf != want) fails++;
    }
    return fails;
}
