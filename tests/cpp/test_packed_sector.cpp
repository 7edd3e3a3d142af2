// Host-side check of the packed ("dense") sector arithmetic (rust_lapper_b200/csrc/packed_sector.cuh):
// fill every sector of a table over random interval sets, then answer queries from single sectors and
// compare with Lapper::count's closed form (src/lib.rs:610-618) evaluated by brute force.  Runs without a
// GPU -- the same functions are compiled into the CUDA kernels.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../rust_lapper_b200/csrc/packed_sector.cuh"

using namespace lap;

static uint64_t brute_count(const std::vector<uint32_t> &S, const std::vector<uint32_t> &E, uint32_t a, uint32_t b) {
    // #starts < b  -  #stops <= a   (wrapping); a == u32::MAX is excluded by packed_locate
    uint64_t lb = std::lower_bound(S.begin(), S.end(), b) - S.begin();
    uint64_t ub = std::upper_bound(E.begin(), E.end(), a) - E.begin();
    return lb - ub;
}

int main() {
    std::mt19937_64 rng(12345);
    uint64_t checked = 0, answered = 0, overflowed = 0, sectors = 0;
    const uint32_t widths[] = {2, 3, 7, 16, 100, 255, 256, 257, 1000, 1550, 1819, 1820};
    for (int round = 0; round < 60; round++) {
        const uint32_t w = widths[round % 12];
        const uint32_t universe_choices[] = {50, 5000, 1000000, 0xFFFFFFFFu};
        const uint32_t U = universe_choices[(round / 12) % 4];
        // about 2.5 starts per bucket on average, clustered so that some buckets overflow and many are empty
        uint64_t n = std::min<uint64_t>(20000, std::max<uint64_t>(1, (uint64_t)((double)U / w * 2.5)));
        if (U == 0xFFFFFFFFu) n = 20000;
        std::vector<uint32_t> S(n), E(n);
        const uint32_t max_len = (round % 3 == 0) ? 0 : (round % 3 == 1 ? w / 2 + 1 : 4 * w + 5);
        for (uint64_t i = 0; i < n; i++) {
            uint32_t s = (uint32_t)(rng() % ((uint64_t)U + 1));
            if (U == 0xFFFFFFFFu && (i & 1)) s = 0xFFFFFFFFu - (uint32_t)(rng() % 100000);  // pile up at the top end
            if (U == 0xFFFFFFFFu && !(i & 1)) s = (uint32_t)(rng() % 100000);               // and at the bottom
            const uint64_t len = max_len ? rng() % ((uint64_t)max_len + 1) : 0;
            const uint64_t e = std::min<uint64_t>((uint64_t)s + len, 0xFFFFFFFFull);
            S[i] = s;
            E[i] = (uint32_t)e;
        }
        std::sort(S.begin(), S.end());
        std::sort(E.begin(), E.end());
        const uint32_t max_coord = std::max(S.back(), E.back());
        // margins: the default w / 8, none, most of the bucket, more than fits (clamped to 2047 - w)
        const uint32_t margins[] = {w / 8, 0, (3 * w) / 4, 160, 5000};
        const PackedParams p = packed_params(w, max_coord, margins[(round / 3) % 5]);
        if (p.w + p.m > 2047 || (uint64_t)p.magic * w < (1ull << 32) || (uint64_t)(p.magic - 1) * w >= (1ull << 32)) {
            printf("bad params for w=%u\n", w);
            return 1;
        }
        // sparse table: only fill the sectors that get queried (a table over 2^32 / 2 coordinates would be huge)
        const int nq = 4000;
        for (int j = 0; j < nq; j++) {
            uint32_t a, b;
            const int kind = j % 8;
            const uint32_t anchor = (kind < 5) ? S[rng() % n] : (uint32_t)(rng() % ((uint64_t)max_coord + 2000));
            const uint32_t qlen = (kind == 0) ? 0 : (kind == 1 ? 1 : (uint32_t)(rng() % (2 * w + 3)));
            const uint32_t back = (uint32_t)(rng() % (w + p.m + 2));
            a = anchor >= back ? anchor - back : 0;
            b = (uint32_t)std::min<uint64_t>((uint64_t)a + qlen, 0xFFFFFFFFull);
            if (kind == 7) std::swap(a, b);                       // inverted query
            if (j == 17) a = 0xFFFFFFFFu, b = 0xFFFFFFFFu;        // the start + 1 wrap
            if (j == 18) a = 0, b = 0;
            if (j == 19) a = 0xFFFFFFFEu, b = 0xFFFFFFFFu;
            if (j >= 20 && j < 40) a = 0xFFFFFFFFu - (uint32_t)(rng() % (p.m + 3)), b = (uint32_t)(rng() % (w + 1));  // a - base + m wraps
            // a start near ZERO with a stop in the last bucket below 2^32: that bucket crosses 2^32 and `a - base + m`
            // wraps into range (count(0, u32::MAX) over intervals that reach the top of the range; found by the fuzz)
            if (j >= 40 && j < 70) a = (uint32_t)(rng() % (2 * w + 2)), b = 0xFFFFFFFFu - (uint32_t)(rng() % (w + 1));
            uint32_t B, xs, te;
            const bool ok = packed_locate(p, a, b, B, xs, te);
            checked++;
            // the bucket must be the floor division clamped to the sentinel
            const uint32_t want_B = std::min<uint32_t>(b / w, p.nb);
            if (B != want_B) {
                printf("bucket of %u at w=%u: got %u want %u\n", b, w, B, want_B);
                return 1;
            }
            if (!ok) continue;
            uint32_t sec[8];
            sectors++;
            if (!packed_fill(S.data(), E.data(), (uint32_t)n, p, B, sec)) {
                overflowed++;
                if (!packed_overflow(sec)) { printf("overflow marker missing\n"); return 1; }
                continue;
            }
            if (packed_overflow(sec)) { printf("spurious overflow marker\n"); return 1; }
            const uint32_t got = packed_count(sec, xs, te);
            const uint32_t want = (uint32_t)brute_count(S, E, a, b);
            answered++;
            if (got != want) {
                printf("round %d w=%u U=%u query (%u, %u) bucket %u: got %u want %u\n", round, w, U, a, b, B, got, want);
                return 1;
            }
        }
    }
    printf("packed sector ok: %llu queries located, %llu answered from one sector, %llu of %llu sectors overflowed\n",
           (unsigned long long)checked, (unsigned long long)answered, (unsigned long long)overflowed,
           (unsigned long long)sectors);
    if (answered < checked / 4) { printf("too few queries exercised the sector path\n"); return 1; }
    return 0;
}
