// query_views.cuh -- what the query kernels share: views of the device index (passed by value to the kernels) and
// the per-query BITS sector arithmetic (index.cuh describes the sector format).  Included by query.cu and
// count_sorted.cu; everything has internal linkage.
#pragma once

#include "index.cuh"

namespace lap {

// what a kernel needs to take the batch-shape decision (see batch_mode below)
struct BatchProbe {
    const uint32_t *qs, *qe;
    uint64_t ntiles;      // full tiles of kProbeTile queries in the batch
    uint32_t n, shift, nb, local_span;
    uint32_t accept;      // bit m set: the kernel runs for batches of mode m (kAcceptAll: no probing)
    uint32_t sorted_ok;   // kBatchSortedDense is on offer
};

// count_sorted.cu: the kernel for sorted dense batches, queries [0, nruns * 256); tile_sums != nullptr: find mode
bool count_sorted_disabled();  // LAPPER_NO_SORTED=1 (experiments / tests)
void launch_count_sorted(const DeviceIndex &ix, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nruns, uint32_t *d_out32,
                         uint64_t *d_out64, uint64_t *tile_sums, const BatchProbe &probe, cudaStream_t stream);

namespace {

struct BucketView {
    const Sector *buckets;
    const uint32_t *S;
    const uint32_t *E;
    uint32_t n;
    const Sector *pool;
    uint32_t shift;
    uint32_t margin;
    uint32_t nb;  // index of the sentinel bucket
    // packed table in front of the sector table (index.cuh); dense == nullptr when there is none
    const Sector *dense;
    PackedParams pk;
    uint32_t local_span;  // a warp whose threads each see their queries within this span stays on the sector-table path
};

// bytes 1 and 3 of t0 and of t1 (the high byte of every 16-bit lane), each replaced by its sign bit
// replicated: 0xFF where the lane's guard bit survived the subtraction, 0x00 where it did not
__device__ __forceinline__ uint32_t guard_bytes(uint32_t t0, uint32_t t1) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(r) : "r"(t0), "r"(t1));
    return r;
}

// acc - #{16-bit lanes of w[2..7] whose value is >= x}; x2 = x replicated in both halves.
// One packed subtract per word, one PRMT per two words, one IDP.4A (bytes are 0 or -1) per PRMT.
__device__ __forceinline__ int sub_lanes_ge(const Sector &s, uint32_t x2, int acc) {
    acc = __dp4a((int)guard_bytes(s.w[2] - x2, s.w[3] - x2), 0x01010101, acc);
    acc = __dp4a((int)guard_bytes(s.w[4] - x2, s.w[5] - x2), 0x01010101, acc);
    acc = __dp4a((int)guard_bytes(s.w[6] - x2, s.w[7] - x2), 0x01010101, acc);
    return acc;
}

__device__ __forceinline__ bool sector_overflow(const Sector &s) { return s.w[7] == 0xFFFFFFFFu; }

// lower_bound / upper_bound inside one overflowing bucket's slice of a sorted array: bisect down to a
// short range, then count it with independent (not dependent) loads -- one memory round trip
__device__ __forceinline__ uint32_t count_lt_range(const uint32_t *__restrict__ a, uint32_t lo, uint32_t hi,
                                                   uint32_t key) {
    while (hi - lo > 32) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    uint32_t c = lo;
#pragma unroll 8
    for (uint32_t i = lo; i < hi; i++) c += (__ldg(a + i) < key);
    return c;
}
__device__ __forceinline__ uint32_t count_le_range(const uint32_t *__restrict__ a, uint32_t lo, uint32_t hi,
                                                   uint32_t key) {
    while (hi - lo > 32) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) <= key)
            lo = mid + 1;
        else
            hi = mid;
    }
    uint32_t c = lo;
#pragma unroll 8
    for (uint32_t i = lo; i < hi; i++) c += (__ldg(a + i) <= key);
    return c;
}

// lane counts of a (non-overflowing) sector of width w
__device__ __forceinline__ uint32_t sector_lb(const Sector &s, uint32_t b_rel) {
    return s.w[0] + kBucketLanes + (uint32_t)sub_lanes_ge(s, b_rel * 0x10001u, 0);
}
__device__ __forceinline__ uint32_t sector_ub(const Sector &s, uint32_t w, uint32_t a_rel_plus_margin) {
    return s.w[1] + (uint32_t)sub_lanes_ge(s, (w + a_rel_plus_margin + 1u) * 0x10001u, 0);
}

// lower_bound(starts, key) from the sector of key's own bucket, following the refinement if any
__device__ __forceinline__ uint32_t lb_from_sector(const BucketView &v, const Sector &s, uint32_t key) {
    const uint32_t low = key & ((1u << v.shift) - 1u);
    if (__builtin_expect(!sector_overflow(s), 1)) return sector_lb(s, low);
    if (s.w[2] != kNoChild) {
        const uint32_t cshift = v.shift - s.w[3];
        const Sector c = ld_sector_cached(v.pool + s.w[2] + (low >> cshift));
        if (!sector_overflow(c)) return sector_lb(c, low & ((1u << cshift) - 1u));
        return count_lt_range(v.S, c.w[0], c.w[4], key);
    }
    return count_lt_range(v.S, s.w[0], s.w[4], key);
}

// upper_bound(stops_sorted, key) from the sector of key's own bucket
__device__ __forceinline__ uint32_t ub_from_sector(const BucketView &v, const Sector &s, uint32_t key) {
    const uint32_t w = 1u << v.shift, low = key & (w - 1u);
    if (__builtin_expect(!sector_overflow(s), 1)) return sector_ub(s, w, low + v.margin);
    if (s.w[2] != kNoChild) {
        const uint32_t cshift = v.shift - s.w[3];
        const Sector c = ld_sector_cached(v.pool + s.w[2] + (low >> cshift));
        if (!sector_overflow(c)) return sector_ub(c, 1u << cshift, low & ((1u << cshift) - 1u));
        return count_le_range(v.E, c.w[5], c.w[1], key);
    }
    return count_le_range(v.E, s.w[5], s.w[1], key);
}

__device__ __forceinline__ uint32_t bucket_of(const BucketView &v, uint32_t x) { return min(x >> v.shift, v.nb); }

// Lapper::count for one query, any case: each side from its own bucket
__device__ __forceinline__ void bits_count_slow(const BucketView &v, uint32_t a, uint32_t b, uint32_t &lb,
                                                uint32_t &ub) {
    const uint32_t bb = bucket_of(v, b), ba = bucket_of(v, a);
    const Sector sb = ld_sector_cached(v.buckets + bb);
    lb = lb_from_sector(v, sb, b);
    if (ba == bb) {
        ub = ub_from_sector(v, sb, a);
    } else {
        const Sector sa = ld_sector_cached(v.buckets + ba);
        ub = ub_from_sector(v, sa, a);
    }
    // src/lib.rs:614: `start + one` wraps to 0 when start == u32::MAX, so `first` is 0 there.
    if (a == 0xFFFFFFFFu) ub = 0;
}

template <bool kOut64>
__device__ __forceinline__ void store_count(uint32_t *o32, uint64_t *o64, uint64_t j, uint32_t lb, uint32_t ub) {
    // len - first - (len - last) in wrapping usize == last - first (mod 2^64)
    if (kOut64)
        st_stream_u64(o64 + j, (uint64_t)lb - (uint64_t)ub);
    else
        st_stream_u32(o32 + j, lb - ub);
}

// ------------------------------------------------------------------------------------- find
// find(a, b) = ascending i with start_i < b and stop_i > a (src/lib.rs:704-717; the start offset
// :632-635 and the early break :712-713 only prune).  The reference scans one window
// [lower_bound(a - max_len), lower_bound(b)) of the whole sorted vector, which a single long interval
// stretches to everything before b (src/lib.rs:36-47).  Here the sorted vector is also kept split by
// length class (index.cuh): per class the window is [lower_bound_c(a - maxlen_c), lower_bound_c(b)),
// the class hit streams are ascending in global position, and a per-query k-way merge emits them in
// exactly the reference's order.
struct FindView {
    BucketView bv;
    bool has_buckets;
    const uint32_t *S;
    const uint32_t *Eiv;  // stops paired with S
    const uint32_t *E;    // sorted stops
    uint32_t n;
    uint32_t max_len;
    bool has_inverted;
    // length classes
    uint32_t ncls;
    uint32_t cls_off[4], cls_n[4], cls_maxlen[4];
    const uint32_t *cs, *ce, *cg;
    const uint4 *dir;
    uint32_t cls_shift, cls_nb;
    // heavy queries
    const uint32_t *pmax, *blkmax32, *blkmax1024;
};

__device__ __forceinline__ uint32_t rank_lb_starts(const FindView &f, uint32_t key) {
    if (f.has_buckets) {
        const uint32_t bk = bucket_of(f.bv, key);
        const Sector s = ld_sector_cached(f.bv.buckets + bk);
        return lb_from_sector(f.bv, s, key);
    }
    return lower_bound_u32(f.S, 0, f.n, key);
}
__device__ __forceinline__ uint32_t rank_ub_stops(const FindView &f, uint32_t key) {
    if (f.has_buckets) {
        const uint32_t bk = bucket_of(f.bv, key);
        const Sector s = ld_sector_cached(f.bv.buckets + bk);
        return ub_from_sector(f.bv, s, key);
    }
    return upper_bound_u32(f.E, 0, f.n, key);
}

__device__ __forceinline__ uint32_t dir_get(const uint4 &d, int c) { return c == 0 ? d.x : c == 1 ? d.y : c == 2 ? d.z : d.w; }

// lower_bound of `key` among the starts of class c: directory rank, then a short forward walk
__device__ __forceinline__ uint32_t class_lb(const FindView &f, int c, uint32_t key) {
    const uint32_t B = min(key >> f.cls_shift, f.cls_nb);
    const uint4 d0 = __ldg(f.dir + B);
    uint32_t r = dir_get(d0, c);
    if (B < f.cls_nb) {
        const uint32_t hi = dir_get(__ldg(f.dir + B + 1), c);
        const uint32_t *cs = f.cs + f.cls_off[c];
        if (hi - r > 16) return lower_bound_u32(cs, r, hi, key);
        while (r < hi && __ldg(cs + r) < key) r++;
    }
    return r;
}

// per-class scan windows of one query; returns the class-relative [lo, hi)
__device__ __forceinline__ void class_window(const FindView &f, int c, uint32_t a, uint32_t b, uint32_t &lo, uint32_t &hi) {
    hi = class_lb(f, c, b);
    lo = class_lb(f, c, a >= f.cls_maxlen[c] ? a - f.cls_maxlen[c] : 0u);
    if (lo > hi) lo = hi;
}

// number of positions find(a, b) yields, by the predicate (any input)
__device__ __forceinline__ uint32_t find_count_scan(const FindView &f, uint32_t a, uint32_t b) {
    uint32_t total = 0;
    for (int c = 0; c < (int)f.ncls; c++) {
        uint32_t lo, hi;
        class_window(f, c, a, b, lo, hi);
        const uint32_t *ce = f.ce + f.cls_off[c];
        for (uint32_t i = lo; i < hi; i++) total += (__ldg(ce + i) > a);
    }
    return total;
}

__device__ __forceinline__ uint32_t find_count_one(const FindView &f, uint32_t a, uint32_t b) {
    if (f.n == 0) return 0;
    if (a < b && !f.has_inverted) {
        // |find| == count when every interval has start <= stop and the query is non-empty
        // (SURVEY Appendix A.4); a == u32::MAX cannot happen here because a < b.
        return rank_lb_starts(f, b) - rank_ub_stops(f, a);
    }
    return find_count_scan(f, a, b);
}


// ---- which kernel serves a large batch is decided ON THE DEVICE, by every warp of every kernel of the pipeline from
// the same 32 sampled places of the batch (no host round trip, no extra pass over the queries): the kernels that
// are not meant for the batch's shape exit at once.
//   kBatchRandom      order-less queries: packed-table path (3 CTAs/SM build of count_bucket_kernel)
//   kBatchLocal       neighbouring queries fall into the same few buckets, but a run of 256 of them spans many
//                     intervals: sector-table path (4 CTAs/SM build)
//   kBatchSortedDense both coordinates non-decreasing and a run of 256 queries spans only a few dozen intervals:
//                     count_sorted_kernel (count_sorted.cu) ranks the run against the sorted arrays directly
constexpr int kBatchRandom = 0, kBatchLocal = 1, kBatchSortedDense = 2;
constexpr uint32_t kAcceptAll = 7u;
constexpr uint32_t kProbeRun = 256;       // queries per sampled run (= the chunk of count_sorted_kernel)
constexpr uint32_t kProbeTile = 1024;     // sampled runs start at multiples of this (the count kernel's tile)
constexpr uint32_t kSortedMaxExpected = 40;  // expected starts (stops) per run for kBatchSortedDense
__device__ __forceinline__ int batch_mode(const BatchProbe &p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t j = (p.ntiles * lane / 32) * kProbeTile + 8 * lane;  // one run per lane
    const uint4 e0 = __ldg(reinterpret_cast<const uint4 *>(p.qe + j)), e1 = __ldg(reinterpret_cast<const uint4 *>(p.qe + j + 4));
    const uint32_t d = e0.w - e0.x;
    const bool local = d + p.local_span < 2u * p.local_span;
    bool dense = false;
    if (p.sorted_ok) {
        const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(p.qs + j)), s1 = __ldg(reinterpret_cast<const uint4 *>(p.qs + j + 4));
        const uint32_t e_last = __ldg(p.qe + j + kProbeRun - 1 - 8 * lane), s_last = __ldg(p.qs + j + kProbeRun - 1 - 8 * lane);
        const uint32_t e_first = __ldg(p.qe + j - 8 * lane), s_first = __ldg(p.qs + j - 8 * lane);
        const bool mono = e0.x <= e0.y && e0.y <= e0.z && e0.z <= e0.w && e0.w <= e1.x && e1.x <= e1.y && e1.y <= e1.z &&
                          e1.z <= e1.w && s0.x <= s0.y && s0.y <= s0.z && s0.z <= s0.w && s0.w <= s1.x && s1.x <= s1.y &&
                          s1.y <= s1.z && s1.z <= s1.w && e_first <= e0.x && e1.w <= e_last && s_first <= s0.x && s1.w <= s_last;
        // expected keys inside the run's span: span * n / (covered coordinate range)
        const uint64_t range = ((uint64_t)p.nb + 1) << p.shift;
        const uint64_t span = (uint64_t)max(e_last - e_first, s_last - s_first);
        dense = mono && span * p.n <= (uint64_t)kSortedMaxExpected * range;
    }
    if (__popc(__ballot_sync(0xFFFFFFFFu, dense)) >= 28) return kBatchSortedDense;
    return __popc(__ballot_sync(0xFFFFFFFFu, local)) >= 28 ? kBatchLocal : kBatchRandom;
}
__device__ __forceinline__ bool batch_accepted(const BatchProbe &p) {
    return p.accept == kAcceptAll || ((p.accept >> batch_mode(p)) & 1u);
}

BucketView make_bucket_view(const DeviceIndex &ix) {
    BucketView v;
    v.buckets = ix.buckets;
    v.pool = ix.pool;
    v.margin = ix.margin;
    v.S = ix.starts;
    v.E = ix.stops_sorted;
    v.n = (uint32_t)ix.n;
    v.shift = ix.shift;
    v.nb = (uint32_t)ix.nbuckets;
    v.dense = ix.dense;
    v.pk.w = ix.dense_w;
    v.pk.m = ix.dense_m;
    v.pk.magic = ix.dense_magic;
    v.pk.nb = (uint32_t)ix.dense_nb;
    v.local_span = ix.shift <= kMaxBucketShift ? 4u << ix.shift : 0u;  // (no sector table: no packed table either)
    // dense sets: the packed table's margin covers longer queries than the sector table's and fewer of its sectors
    // overflow, so it serves start-sorted batches better too -- every warp takes the packed path
    if (ix.dense && ix.dense_m > ix.margin) v.local_span = 0;
    return v;
}

FindView make_find_view(const DeviceIndex &ix) {
    FindView f;
    f.has_buckets = ix.buckets != nullptr;
    f.bv = make_bucket_view(ix);
    f.S = ix.starts;
    f.Eiv = ix.stops_by_iv;
    f.E = ix.stops_sorted;
    f.n = (uint32_t)ix.n;
    f.max_len = ix.max_len;
    f.has_inverted = ix.has_inverted;
    f.ncls = ix.ncls;
    for (int c = 0; c < 4; c++) {
        f.cls_off[c] = ix.cls_off[c];
        f.cls_n[c] = ix.cls_n[c];
        f.cls_maxlen[c] = ix.cls_maxlen[c];
    }
    f.cs = ix.cls_starts;
    f.ce = ix.cls_stops;
    f.cg = ix.cls_gidx;
    f.dir = ix.cls_dir;
    f.cls_shift = ix.cls_shift;
    f.cls_nb = (uint32_t)ix.cls_nb;
    f.pmax = ix.prefix_max;
    f.blkmax32 = ix.blkmax32;
    f.blkmax1024 = ix.blkmax1024;
    return f;
}

uint32_t capped_grid(uint64_t work_items, int per_block, int waves) {
    uint64_t g = div_up(work_items ? work_items : 1, per_block);
    const uint64_t cap = (uint64_t)sm_count() * waves;
    return (uint32_t)(g < cap ? g : cap);
}


}  // namespace
}  // namespace lap
