// packed_sector.cuh -- the packed ("dense") sector: format, build and query arithmetic.
//
// Host + device: the same functions fill a sector (build.cu), answer a query from it (query.cu) and are
// exercised on the CPU against a brute-force count by tests/cpp/test_packed_sector.cpp (the only
// device-specific piece is the 224-bit borrow chain, written with sub.cc/subc.cc on the device).
//
// Why a second table: the BITS sector table (index.cuh) is read at one random 32-byte sector per query, so
// for query batches without locality its size decides how much of it stays in L2 (tools/mb_gather.cu: the
// gather rate falls from 286 G sectors/s to 166 G/s between a 64 MB and a 97 MB table).  The packed table
// holds the same keys in 18 twelve-bit lanes per sector over buckets of ANY width w (not a power of two;
// w + m <= 2047), which makes it about 1.5x smaller.  It answers only the common case (query start within
// the margin below the bucket of the query stop, bucket not overflowing); everything else falls back to
// the sector table.
//
// Bucket B = floor(x / w) covers [base, base + w), base = B * w, stop margin m (build.cu::choose_dense_shape: w / 8, but
// at least ~160 coordinates where the buckets are narrow, so that a read-sized query needs one sector):
//   w[0]      h = #starts < base  -  #sorted stops < base - m  +  L          (wrapping u32)
//   w[1..7]   a 224-bit little-endian number: lane i (12 bits: guard bit 11 set, 11 value bits) at bit 12 i,
//             i < 18; bits 216..223 are flags.  Lanes [0, L) are the S region, lanes [L, 18) the E region,
//             L in {8, 9, 10} chosen per sector (flag bit 24 of w[7]: L > 8, bit 25: L > 9):
//               S lane: a start s in [base, base + w)     ->  s - base               empty -> 0x7FF
//               E lane: a stop  e in [base - m, base + w) ->  2047 - (e - base + m)  empty -> 0
//             bit 31 of w[7]: the bucket does not fit (more than L starts or 18 - L stops): fall back.
//   Lapper::count(a, b) (src/lib.rs:610-618) for b in the bucket and a in [base - m, base + w):
//       #starts < b   = #starts < base + (#S lanes) - #{S lane >= b - base}       (empty S lanes always count)
//       #stops  <= a  = #stops < base - m + #{E lane >= 2047 - (a - base + m)}    (empty E lanes never count)
//       count = h - popc(guard bits that survive ONE 224-bit subtraction of the per-lane thresholds)
//   (lane - threshold never borrows out of a lane because of the guard bit, so the 7-word borrow chain only
//   serves lanes that straddle two words.)
//   Sector nb is a sentinel for coordinates beyond the last key: no starts of its own, only its margin's stops.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define LAPPER_HD __host__ __device__ __forceinline__
#else
#define LAPPER_HD static inline
#endif

namespace lap {

constexpr uint32_t kDenseLanes = 18;
constexpr uint32_t kDenseMaxWidth = 1820;  // w + w / 8 <= 2047 (margins are clamped to 2047 - w)
constexpr uint32_t kDenseMinWidth = 16;
constexpr uint32_t kDenseOverflow = 0x80000000u;

struct PackedParams {
    uint32_t w;      // bucket width
    uint32_t m;      // stop margin
    uint32_t magic;  // ceil(2^32 / w)
    uint32_t nb;     // index of the sentinel sector = floor(max_coord / w) + 1
};

LAPPER_HD uint32_t pk_min(uint32_t a, uint32_t b) { return a < b ? a : b; }
LAPPER_HD uint32_t pk_max(uint32_t a, uint32_t b) { return a > b ? a : b; }

// m: stop margin, clamped so that w + m <= 2047 (lane values have 11 bits)
LAPPER_HD PackedParams packed_params(uint32_t w, uint32_t max_coord, uint32_t m) {
    PackedParams p;
    p.w = w;
    p.m = pk_min(m, 2047u - w);
    p.magic = (uint32_t)(((1ull << 32) + w - 1) / w);
    p.nb = max_coord / w + 1;
    return p;
}

LAPPER_HD uint32_t pk_umulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
LAPPER_HD uint32_t pk_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}

// #elements of the sorted array a[lo, hi) that are < key (key may exceed 32 bits)
LAPPER_HD uint32_t pk_lower_bound(const uint32_t *a, uint32_t lo, uint32_t hi, uint64_t key) {
    if (key > 0xFFFFFFFFull) return hi;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < (uint32_t)key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// Sector of the query stop b, the S-lane threshold xs = b - base and the E-lane threshold
// te = 2047 - (a - base + m).  False when that sector cannot answer the query: a is not within the margin
// below the bucket (long query; every inverted query), or a == u32::MAX (the `start + 1` wrap of src/lib.rs:614).
LAPPER_HD bool packed_locate(const PackedParams &p, uint32_t a, uint32_t b, uint32_t &B, uint32_t &xs, uint32_t &te) {
    uint32_t q = pk_umulhi(b, p.magic);  // floor(b / w) or one more
    if (b - q * p.w >= p.w) q--;         // remainder wrapped: one too many
    B = pk_min(q, p.nb);                 // beyond the last key: the sentinel
    const uint32_t base = B * p.w;
    xs = pk_min(b - base, 2047u);        // only the sentinel can see more than w; it has no S keys
    const uint32_t a_rel = a - base + p.m;
    te = (2047u - a_rel) & 0x7FFu;
    // a <= b: an inverted query whose start lies within m of 2^32 would wrap a_rel into range for bucket 0
    return a_rel < p.w + p.m && a <= b && a != 0xFFFFFFFFu;
}

// twelve-bit replication of x over the three word phases of a 96-bit period (lane i at bit 12 i)
LAPPER_HD uint32_t rep12_p0(uint32_t x) { return x * 0x01001001u; }
LAPPER_HD uint32_t rep12_p1(uint32_t x) { return x * 0x10010010u + (x >> 8); }
LAPPER_HD uint32_t rep12_p2(uint32_t x) { return x * 0x00100100u + (x >> 4); }

// Lapper::count from one packed sector (not overflowing, query located by packed_locate)
LAPPER_HD uint32_t packed_count(const uint32_t s[8], uint32_t xs, uint32_t te) {
    // lanes 8 and 9 belong to the S region when L > 8 / L > 9 (flag bits 24 / 25 of w[7])
    const uint32_t m8 = (uint32_t)((int32_t)(s[7] << 7) >> 31), m9 = (uint32_t)((int32_t)(s[7] << 6) >> 31);
    const uint32_t d = xs ^ te;
    const uint32_t t[7] = {rep12_p0(xs), rep12_p1(xs), rep12_p2(xs),
                           (te ^ (d & m8)) | ((te ^ (d & m9)) << 12) | (te << 24),
                           rep12_p1(te), rep12_p2(te), rep12_p0(te)};
    uint32_t r[7];
#if defined(__CUDA_ARCH__)
    asm("sub.cc.u32 %0, %7, %14;\n\t"
        "subc.cc.u32 %1, %8, %15;\n\t"
        "subc.cc.u32 %2, %9, %16;\n\t"
        "subc.cc.u32 %3, %10, %17;\n\t"
        "subc.cc.u32 %4, %11, %18;\n\t"
        "subc.cc.u32 %5, %12, %19;\n\t"
        "subc.u32 %6, %13, %20;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6])
        : "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]), "r"(t[0]), "r"(t[1]), "r"(t[2]),
          "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]));
#else
    uint32_t borrow = 0;
    for (int j = 0; j < 7; j++) {
        const uint64_t x = (uint64_t)s[1 + j] - (uint64_t)t[j] - borrow;
        r[j] = (uint32_t)x;
        borrow = (uint32_t)(x >> 63);
    }
#endif
    // guard bits: bit 12 i + 11 of the 224-bit number = bits {11, 23} / {3, 15, 27} / {7, 19, 31} of the three
    // words of a 96-bit group -- disjoint positions, so a group ORs into one word with a guard every 4 bits
    const uint32_t g0 = (r[0] & 0x00800800u) | (r[1] & 0x08008008u) | (r[2] & 0x80080080u);
    const uint32_t g1 = (r[3] & 0x00800800u) | (r[4] & 0x08008008u) | (r[5] & 0x80080080u);
    const uint32_t g2 = r[6] & 0x00800800u;
    return s[0] - pk_popc(g0 | (g1 >> 1) | (g2 >> 2));
}

LAPPER_HD bool packed_overflow(const uint32_t s[8]) { return (s[7] >> 31) != 0; }

// Sector B of the packed table over the sorted starts S[n] and the sorted stops E[n].  Returns false
// (and writes an overflow marker) when the bucket's keys do not fit.
LAPPER_HD bool packed_fill(const uint32_t *S, const uint32_t *E, uint32_t n, const PackedParams &p, uint64_t B,
                           uint32_t out[8]) {
    const uint64_t base = B * p.w, next = base + p.w;
    const uint64_t lo_m = base >= p.m ? base - p.m : 0;
    const uint32_t rs = pk_lower_bound(S, 0, n, base), rs_next = pk_lower_bound(S, rs, n, next);
    const uint32_t re_m = pk_lower_bound(E, 0, n, lo_m), re_hi = pk_lower_bound(E, re_m, n, next);
    const uint32_t ns = rs_next - rs, ne = re_hi - re_m;
    const uint32_t L = pk_min(pk_max(ns, 8u), 10u);
    for (int j = 0; j < 8; j++) out[j] = 0;
    // The bucket that crosses 2^32 (base + w > 2^32: only the last one can) is never answered from here: for a query
    // that starts within base + w - 2^32 of ZERO and stops in this bucket, packed_locate's `a - base + m` wraps into
    // range as if the start lay inside the bucket.  Marked as overflowing, its queries go to the sector table
    // (found by tests/fuzz_parity.py: count(0, u32::MAX) over intervals that reach the top of the range).
    if (ns > L || ne > kDenseLanes - L || next > (1ull << 32)) {
        out[7] = kDenseOverflow;
        return false;
    }
    out[0] = rs - re_m + L;
    uint32_t words[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // words[0..6] = the 224-bit lane number, words[7] catches the spill
    for (uint32_t i = 0; i < kDenseLanes; i++) {
        uint32_t v;
        if (i < L)
            v = i < ns ? (uint32_t)((uint64_t)S[rs + i] - base) : 0x7FFu;
        else
            v = (i - L) < ne ? 2047u - (uint32_t)((uint64_t)E[re_m + (i - L)] + p.m - base) : 0u;
        v |= 0x800u;
        const uint32_t bit = 12u * i, k = bit >> 5, sh = bit & 31u;
        words[k] |= v << sh;
        if (sh > 20u) words[k + 1] |= v >> (32u - sh);
    }
    for (int j = 0; j < 7; j++) out[1 + j] = words[j];
    out[7] |= (L > 8u ? 1u << 24 : 0u) | (L > 9u ? 1u << 25 : 0u);
    return true;
}

}  // namespace lap
