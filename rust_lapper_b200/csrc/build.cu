// build.cu -- index build: Lapper::new (/root/reference/src/lib.rs:196-236) on the device.
//
//   sort by (start, stop), stable   -> hand-written LSD radix sort (8-bit digits) of the 64-bit key
//                                      (start << 32) | stop with the input position as payload
//                                      (src/lib.rs:145-157, :201); stable like Vec::sort
//   starts / stops                  -> unzip of the sorted keys; stops sorted on their own (:203-216)
//   max_len                         -> max(stop (-) start) reduction, checked_sub -> 0 (:218-227)
//   + prefix_max, + the BITS bucket sectors described in index.cuh (new; the reference searches
//     the sorted vectors directly).
#include <algorithm>
#include <cstdlib>

#include "index.cuh"

namespace lap {

namespace {

constexpr int kThreads = 256;

__global__ void pack_keys_kernel(const uint32_t *__restrict__ s, const uint32_t *__restrict__ e, uint64_t n,
                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ pos) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = ((uint64_t)s[i] << 32) | (uint64_t)e[i];
    pos[i] = (uint32_t)i;
}

__global__ void unpack_keys_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t *__restrict__ s,
                                   uint32_t *__restrict__ e) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = keys[i];
    s[i] = (uint32_t)(k >> 32);
    e[i] = (uint32_t)k;
}

// ---- LSD radix sort, 8-bit digits.  One pass = histogram kernel -> exclusive scan of the
// (digit-major, block-minor) histogram -> scatter kernel.  A block owns a tile of 4096 consecutive
// items; inside the tile, warp w owns 512 consecutive items as 16 rows of 32, and items are ranked
// within their digit in exactly that order (row by row, lane by lane) with __match_any_sync, so every
// pass is stable and the whole sort keeps equal keys in input order.
constexpr int kRsItems = 16;
constexpr int kRsTile = kThreads * kRsItems;  // 4096
constexpr int kRsWarps = kThreads / 32;

template <typename K>
__device__ __forceinline__ uint32_t rs_digit(K k, int shift) {
    return (uint32_t)(k >> shift) & 255u;
}

template <typename K>
__global__ void __launch_bounds__(kThreads) rs_hist_kernel(const K *__restrict__ keys, uint64_t n, int shift,
                                                           uint32_t *__restrict__ block_hist, uint32_t nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kRsTile;
#pragma unroll
    for (int i = 0; i < kRsItems; i++) {
        const uint64_t idx = base + (uint64_t)i * kThreads + threadIdx.x;
        if (idx < n) atomicAdd(&h[rs_digit(keys[idx], shift)], 1u);
    }
    __syncthreads();
    block_hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

template <typename K, bool kHasVal>
__global__ void __launch_bounds__(kThreads) rs_scatter_kernel(const K *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                              K *__restrict__ kout, uint32_t *__restrict__ vout,
                                                              uint64_t n, int shift,
                                                              const uint32_t *__restrict__ block_off, uint32_t nblocks) {
    __shared__ uint32_t wcount[kRsWarps][256];
    __shared__ uint32_t gbase[256];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kRsWarps * 256; i += kThreads) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const uint64_t seg = (uint64_t)blockIdx.x * kRsTile + (uint64_t)w * (32 * kRsItems);
    K k[kRsItems];
    uint32_t v[kRsItems], rank[kRsItems];
#pragma unroll
    for (int r = 0; r < kRsItems; r++) {
        const uint64_t idx = seg + (uint64_t)r * 32 + lane;
        const bool valid = idx < n;
        k[r] = valid ? kin[idx] : (K)0;
        if (kHasVal) v[r] = valid ? vin[idx] : 0u;
        const uint32_t d = rs_digit(k[r], shift);
        // lanes holding the same digit in this row; out-of-range lanes only match themselves
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, valid ? d : 256u + lane);
        rank[r] = wcount[w][d] + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
        if (valid && lane == (uint32_t)(__ffs(peers) - 1)) wcount[w][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // thread t = digit t: exclusive prefix of the warps' counts, and the block's global base for the digit
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < kRsWarps; i++) {
            const uint32_t c = wcount[i][threadIdx.x];
            wcount[i][threadIdx.x] = acc;
            acc += c;
        }
        gbase[threadIdx.x] = block_off[(uint64_t)threadIdx.x * nblocks + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRsItems; r++) {
        const uint64_t idx = seg + (uint64_t)r * 32 + lane;
        if (idx < n) {
            const uint32_t d = rs_digit(k[r], shift);
            const uint32_t pos = gbase[d] + wcount[w][d] + rank[r];
            kout[pos] = k[r];
            if (kHasVal) vout[pos] = v[r];
        }
    }
}

// Small sets (a per-chromosome or per-contig Lapper of a few thousand intervals is the reference's everyday use): the
// radix sort's 24 launches per 64-bit sort are all overhead there.  Every CTA keeps ALL keys in shared memory and ranks
// 128 of them by counting: rank_i = #{j < i : k_j <= k_i} + #{j > i : k_j < k_i} -- stable, like the radix passes, so
// ties keep input order (src/lib.rs:201: `intervals.sort()` is stable).  Out of place, copied back by the caller.
constexpr uint32_t kSmallSortMax = 4096, kSmallSortThreads = 128;
inline bool small_sort_disabled() {  // LAPPER_NO_SMALL_SORT=1: the radix passes for every size (A/B, tests of the passes on tiny sets)
    static const bool off = [] {
        const char *e = std::getenv("LAPPER_NO_SMALL_SORT");
        return e && *e && *e != '0';
    }();
    return off;
}
template <typename K, bool kHasVal>
__global__ void __launch_bounds__(kSmallSortThreads) small_sort_kernel(const K *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                                      K *__restrict__ kout, uint32_t *__restrict__ vout, uint32_t n) {
    __shared__ K sk[kSmallSortMax];
    for (uint32_t j = threadIdx.x; j < n; j += kSmallSortThreads) sk[j] = kin[j];
    __syncthreads();
    const uint32_t i = blockIdx.x * kSmallSortThreads + threadIdx.x;
    if (i >= n) return;
    const K me = sk[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < i; j++) rank += sk[j] <= me;       // (the lanes of a warp read the same word: a broadcast)
    for (uint32_t j = i + 1; j < n; j++) rank += sk[j] < me;
    kout[rank] = me;
    if (kHasVal) vout[rank] = vin[i];
}

// Sorts n keys (and their u32 payloads) by bits [0, nbits) of the key.  `keys`/`vals` hold the input and,
// on return, the output (nbits is a multiple of 16, so the ping-pong ends where it started).
template <typename K, bool kHasVal>
void radix_sort(K *keys, K *keys_tmp, uint32_t *vals, uint32_t *vals_tmp, uint64_t n, int nbits, cudaStream_t stream) {
    if (n <= kSmallSortMax && nbits == (int)(8 * sizeof(K)) && !small_sort_disabled()) {
        small_sort_kernel<K, kHasVal><<<(uint32_t)div_up(n, kSmallSortThreads), kSmallSortThreads, 0, stream>>>(keys, vals, keys_tmp, vals_tmp, (uint32_t)n);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(keys, keys_tmp, n * sizeof(K), cudaMemcpyDeviceToDevice, stream));
        if (kHasVal) LAPPER_CUDA_CHECK(cudaMemcpyAsync(vals, vals_tmp, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
        return;
    }
    const uint32_t nblocks = (uint32_t)div_up(n, kRsTile);
    Scratch<uint32_t> hist((uint64_t)256 * nblocks, stream);
    K *kin = keys, *kout = keys_tmp;
    uint32_t *vin = vals, *vout = vals_tmp;
    for (int shift = 0; shift < nbits; shift += 8) {
        rs_hist_kernel<K><<<nblocks, kThreads, 0, stream>>>(kin, n, shift, hist.p, nblocks);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        launch_exclusive_sum(hist.p, hist.p, (uint64_t)256 * nblocks, stream);
        rs_scatter_kernel<K, kHasVal><<<nblocks, kThreads, 0, stream>>>(kin, vin, kout, vout, n, shift, hist.p, nblocks);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
}

// One pass over the sorted intervals: max_len (src/lib.rs:218-227), the inverted flag, the histogram of lengths by
// bit length (what the length classes are chosen from), the largest input position, and the end points the table
// shapes depend on.
__global__ void __launch_bounds__(kThreads) stats_kernel(const uint32_t *__restrict__ s, const uint32_t *__restrict__ e,
                                                         const uint32_t *__restrict__ e_sorted,
                                                         const uint32_t *__restrict__ perm, uint64_t n,
                                                         BuildStats *__restrict__ out) {
    __shared__ uint32_t cnt[33];
    __shared__ uint32_t mx[33];
    for (int i = threadIdx.x; i < 33; i += blockDim.x) cnt[i] = 0, mx[i] = 0;
    __syncthreads();
    uint32_t inv = 0, pmax = 0;
    const uint32_t lane = threadIdx.x & 31u;
    // (warp-uniform trip count: the loop runs on the warp's first index)
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < n; i0 += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = i0 + lane;
        const bool live = i < n;
        const uint32_t a = live ? s[i] : 0u, b = live ? e[i] : 0u;
        const uint32_t len = b > a ? b - a : 0u;
        const uint32_t bin = len ? 32u - __clz(len) : 0u;
        // lanes of a warp mostly share a bin: one shared-memory atomic per distinct bin
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, live ? bin : 64u + lane);
        const uint32_t len_max = __reduce_max_sync(peers, len);
        if (live && lane == (uint32_t)(__ffs(peers) - 1)) {
            atomicAdd(&cnt[bin], (uint32_t)__popc(peers));
            atomicMax(&mx[bin], len_max);
        }
        inv |= (b < a);
        if (live) pmax = max(pmax, perm[i]);
    }
    inv = __reduce_or_sync(0xFFFFFFFFu, inv);
    pmax = __reduce_max_sync(0xFFFFFFFFu, pmax);
    if ((threadIdx.x & 31) == 0) {
        if (inv) atomicOr(&out->inverted, 1u);
        if (pmax) atomicMax(&out->max_perm, pmax);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 33; i += blockDim.x) {
        if (cnt[i]) {
            atomicAdd(&out->count[i], (unsigned long long)cnt[i]);
            atomicMax(&out->maxlen[i], mx[i]);
            atomicMax(&out->max_len, mx[i]);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out->first_start = s[0];
        out->last_start = s[n - 1];
        out->last_stop = e_sorted[n - 1];
    }
}

// Lanes of one sector covering [base, base + w) with stop margin m, from the slices
// S[rs, rs_next) and E[re_m, re_hi) of the sorted arrays (re_m = #stops < base - m).
__device__ __forceinline__ void pack_lanes(Sector &sec, const uint32_t *__restrict__ S, const uint32_t *__restrict__ E,
                                           uint64_t base, uint32_t w, uint32_t m, uint32_t rs, uint32_t ns,
                                           uint32_t re_m, uint32_t ne) {
    uint32_t lanes[kBucketLanes];
#pragma unroll
    for (uint32_t j = 0; j < kBucketLanes; j++) {
        uint32_t v = 0x8000u | w;
        if (j < ns)
            v = 0x8000u | (uint32_t)((uint64_t)S[rs + j] - base);
        else if (j < ns + ne)
            v = 0x8000u | (w + (uint32_t)((uint64_t)E[re_m + (j - ns)] + m - base));
        lanes[j] = v;
    }
#pragma unroll
    for (int j = 0; j < 6; j++) sec.w[2 + j] = lanes[2 * j] | (lanes[2 * j + 1] << 16);
}

__device__ __forceinline__ uint32_t lb64(const uint32_t *__restrict__ a, uint32_t lo, uint32_t hi, uint64_t key) {
    return key > 0xFFFFFFFFull ? hi : lower_bound_u32(a, lo, hi, (uint32_t)key);
}

// One thread per bucket: coherent bisections (neighbouring buckets walk the same path) and one
// 32-byte sector store.  Overflowing buckets choose their refinement depth and reserve child sectors.
__global__ void fill_buckets_kernel(const uint32_t *__restrict__ S, const uint32_t *__restrict__ E, uint32_t n,
                                    uint32_t shift, uint32_t margin, uint64_t nb, Sector *__restrict__ buckets,
                                    unsigned long long *__restrict__ pool_count) {
    const uint64_t B = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (B > nb) return;
    Sector sec;
    const uint32_t w = 1u << shift;
    // B == nb is the sentinel (first bucket past the last key): no keys of its own, rs = re_hi = n fall out of
    // the bisections below, but it does hold the stops of its margin like every other bucket
    const uint64_t base = B << shift, next = base + w;
    const uint64_t lo_m = base >= margin ? base - margin : 0;
    const uint32_t rs = lb64(S, 0, n, base), rs_next = lb64(S, rs, n, next);
    const uint32_t re_m = lb64(E, 0, n, lo_m), re_lo = lb64(E, re_m, n, base), re_hi = lb64(E, re_lo, n, next);
    const uint32_t ns = rs_next - rs, ne = re_hi - re_m;
    sec.w[0] = rs;
    sec.w[1] = re_hi;
    if (ns + ne <= kBucketLanes) {
        // base - m may be negative for bucket 0: the lane offset uses the real margin, not the clamped one
        pack_lanes(sec, S, E, base, w, margin, rs, ns, re_m, ne);
        buckets[B] = sec;
        return;
    }
    // overflow: find the smallest depth j whose 2^j children all fit (own keys only, no margin);
    // more children than keys would only waste sectors, so j is capped at ceil(log2(keys))
    const uint32_t own = ns + (re_hi - re_lo);
    uint32_t jmax = 0;
    while ((1u << jmax) < own && jmax < shift) jmax++;
    uint32_t j = 0;
    for (uint32_t t = 1; t <= jmax; t++) {
        const uint32_t cshift = shift - t;
        uint32_t is = rs, ie = re_lo, worst = 0;
        for (uint32_t c = 0; c < (1u << t); c++) {
            const uint64_t cnext = base + ((uint64_t)(c + 1) << cshift);
            uint32_t cnt = 0;
            while (is < rs_next && (uint64_t)S[is] < cnext) is++, cnt++;
            while (ie < re_hi && (uint64_t)E[ie] < cnext) ie++, cnt++;
            worst = max(worst, cnt);
        }
        j = t;
        if (worst <= kBucketLanes) break;
    }
    sec.w[2] = j ? (uint32_t)atomicAdd(pool_count, (unsigned long long)(1u << j)) : kNoChild;
    sec.w[3] = j;
    sec.w[4] = rs_next;
    sec.w[5] = re_lo;
    sec.w[6] = 0;
    sec.w[7] = 0xFFFFFFFFu;
    buckets[B] = sec;
}

// Children of the overflowing buckets (one thread per parent; a parent has 2^j children, j small).
__global__ void fill_children_kernel(const uint32_t *__restrict__ S, const uint32_t *__restrict__ E, uint32_t n,
                                     uint32_t shift, uint64_t nb, const Sector *__restrict__ buckets,
                                     Sector *__restrict__ pool) {
    const uint64_t B = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (B > nb) return;
    const Sector par = buckets[B];
    if (par.w[7] != 0xFFFFFFFFu || par.w[2] == kNoChild) return;
    const uint32_t j = par.w[3], cshift = shift - j, cw = 1u << cshift;
    const uint64_t base = B << shift;
    uint32_t is = par.w[0], ie = par.w[5];
    const uint32_t rs_end = par.w[4], re_end = par.w[1];
    for (uint32_t c = 0; c < (1u << j); c++) {
        const uint64_t cbase = base + ((uint64_t)c << cshift), cnext = cbase + cw;
        const uint32_t rs = is, re_lo = ie;
        while (is < rs_end && (uint64_t)S[is] < cnext) is++;
        while (ie < re_end && (uint64_t)E[ie] < cnext) ie++;
        Sector sec;
        sec.w[0] = rs;
        sec.w[1] = ie;
        const uint32_t ns = is - rs, ne = ie - re_lo;
        if (ns + ne <= kBucketLanes) {
            pack_lanes(sec, S, E, cbase, cw, 0, rs, ns, re_lo, ne);
        } else {  // still too many (equal coordinates): leaf, resolved against the sorted arrays
            sec.w[2] = kNoChild;
            sec.w[3] = 0;
            sec.w[4] = is;
            sec.w[5] = re_lo;
            sec.w[6] = 0;
            sec.w[7] = 0xFFFFFFFFu;
        }
        pool[par.w[2] + c] = sec;
    }
}

// Packed sectors (packed_sector.cuh): one thread per bucket of width w (any width), one 32-byte store.
__global__ void fill_dense_kernel(const uint32_t *__restrict__ S, const uint32_t *__restrict__ E, uint32_t n,
                                  PackedParams p, Sector *__restrict__ dense,
                                  unsigned long long *__restrict__ overflow_count) {
    const uint64_t B = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (B > p.nb) return;  // B == nb: the sentinel
    Sector sec;
    if (!packed_fill(S, E, n, p, B, sec.w)) atomicAdd(overflow_count, 1ull);
    dense[B] = sec;
}

// max stop of every 32 consecutive sorted intervals (one warp per block), then of every 32 such blocks
__global__ void __launch_bounds__(kThreads) block_max_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                             uint32_t *__restrict__ out, uint64_t nblocks) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (warp >= nblocks) return;
    const uint64_t i = warp * 32 + lane;
    const uint32_t v = __reduce_max_sync(0xFFFFFFFFu, i < n ? in[i] : 0u);
    if (lane == 0) out[warp] = v;
}

// rank of each bucket's first coordinate in every class (one uint4 per bucket)
__global__ void fill_class_dir_kernel(const uint32_t *__restrict__ cs, uint32_t off0, uint32_t n0, uint32_t off1,
                                      uint32_t n1, uint32_t off2, uint32_t n2, uint32_t off3, uint32_t n3,
                                      uint32_t shift, uint64_t nb, uint4 *__restrict__ dir) {
    const uint64_t B = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (B > nb) return;
    uint4 r;
    if (B == nb) {
        r = make_uint4(n0, n1, n2, n3);
    } else {
        const uint32_t base = (uint32_t)(B << shift);
        r.x = lower_bound_u32(cs + off0, 0, n0, base);
        r.y = lower_bound_u32(cs + off1, 0, n1, base);
        r.z = lower_bound_u32(cs + off2, 0, n2, base);
        r.w = lower_bound_u32(cs + off3, 0, n3, base);
    }
    dir[B] = r;
}

uint32_t grid_for(uint64_t n) { return (uint32_t)div_up(n ? n : 1, kThreads); }

// Pick the bucket width: about six keys (starts + stops) per 12-lane sector keeps the overflow
// rate of a Poisson load under 2%, and at the cfg-2 shape (10 M intervals over 3.1 G) that is
// shift 10 -> 3.0 M sectors = 97 MB, which fits the 126 MB L2.
bool choose_shift(uint64_t n, uint32_t max_coord, uint32_t flags, uint32_t &shift_out) {
    if (flags & LAPPER_BUILD_NO_BUCKETS) return false;
    const uint32_t forced = (flags >> 8) & 0xFFu;
    if (forced) {
        uint32_t k = forced - 1u;
        LAPPER_REQUIRE(k <= kMaxBucketShift, "LAPPER_BUILD_SHIFT: shift must be <= 13");
        // a forced width never makes the table larger than 2^26 sectors (2 GB); widen it instead
        while (k < kMaxBucketShift && ((uint64_t)(max_coord >> k) + 1) > (1ull << 26)) k++;
        shift_out = k;
        return true;
    }
    const bool force = (flags & LAPPER_BUILD_FORCE_BUCKETS) != 0;
    if (n < 4096 && !force) return false;
    const uint64_t target = (2 * n) / 6 > 1024 ? (2 * n) / 6 : 1024;  // ~6.6 own keys + ~0.8 margin stops per sector
    uint32_t k = 0;
    while (k < kMaxBucketShift && ((uint64_t)(max_coord >> k) + 1) > target) k++;
    const uint64_t nb = (uint64_t)(max_coord >> k) + 1;
    if (!force) {
        if (nb > 64 * n) return false;      // very sparse: the table would dwarf the index
        if (2 * n > 24 * nb) return false;  // very dense: nearly every sector would overflow
    }
    shift_out = k;
    return true;
}

// Shape of the packed table: bucket width w and stop margin m.
//   * fill: about 4.25 starts (and 4.8 stops incl. the margin's) per 18-lane sector, i.e. 7.5 bytes per interval --
//     75 MB at the cfg-2 shape with 1.6 % of the sectors overflowing.  Measured there
//     (profiles/r1_exp_packed_table_fill*.log): 4.0 -> 0.688 ms, 4.25 -> 0.665, 4.5 -> 0.672, 5.0 -> 0.725, 5.5 -> 0.788
//     (a smaller table stays in L2 better, but every overflowing sector sends its queries through the deferred path).
//   * margin: w / 8, but at least kDenseMinMargin coordinates (capped at 3/4 w): a query whose start lies further
//     below its stop's bucket needs a second sector, and on dense sets the buckets are narrower than a read-sized
//     query (50 M intervals: w = 263).  The extra margin stops are paid for by a narrower bucket at the same fill.
//   * where: next to a sector table too large to stay in L2 -- either the packed one does (<= 128 MB and clearly
//     smaller), or neither does and the packed table saves second DRAM accesses (its margin covers longer queries
//     than the sector table's w / 4).
constexpr uint32_t kDenseMinMargin = 160;
constexpr uint32_t kQueryLenHintMax = 640;  // longer queries leave too few lanes for the bucket itself: default margin
bool choose_dense_shape(uint64_t n, uint32_t max_coord, uint32_t flags, uint64_t sector_table_bytes, uint32_t sector_margin,
                        uint32_t query_len_hint, uint32_t &w_out, uint32_t &m_out) {
    if (flags & LAPPER_BUILD_NO_DENSE) return false;
    uint32_t min_margin = kDenseMinMargin;
    // a caller that announced its query length (lapper_tune_query_length) gets a margin that covers it: measured at
    // the cfg-2 shape, random order, per 100 M queries: 250 bp 0.942 -> 0.740 ms, 400 bp 1.345 -> 0.839 ms (and 150 bp
    // 0.688 -> 0.743 / 0.837 ms with those wider margins, which is why the default stays where it is)
    if (query_len_hint && query_len_hint <= kQueryLenHintMax) min_margin = std::max(min_margin, query_len_hint + 20);
    if (const char *env = std::getenv("LAPPER_DENSE_MIN_MARGIN")) min_margin = (uint32_t)atoi(env);  // experiments
    const auto margin_for = [&](uint32_t w) { return std::min(std::max(w / 8, std::min(min_margin, (3 * w) / 4)), 2047u - w); };
    const uint32_t forced = (flags >> 16) & 0xFFFu;
    if (forced) {
        LAPPER_REQUIRE(forced >= 2 && forced <= kDenseMaxWidth, "LAPPER_BUILD_DENSE_WIDTH: width must be in [2, 1820]");
        // a forced width never makes the table larger than 2^26 sectors (2 GB); widen it instead
        uint32_t w = forced;
        while (w < kDenseMaxWidth && (uint64_t)max_coord / w + 2 > (1ull << 26)) w = std::min(2 * w, kDenseMaxWidth);
        w_out = w;
        m_out = w / 8;
        return true;
    }
    const bool force = (flags & LAPPER_BUILD_FORCE_DENSE) != 0;
    double per_sector = 4.25;
    if (const char *env = std::getenv("LAPPER_DENSE_STARTS_PER_SECTOR")) per_sector = std::max(0.5, atof(env));  // experiments
    const double span = (double)max_coord + 1.0;
    const double density = (double)std::max<uint64_t>(n, 1) / span;  // intervals per coordinate
    const double entries = per_sector * 2.125;                      // lanes in use per sector at margin w / 8
    double w = per_sector / density;
    for (int it = 0; it < 4; it++) {  // narrower buckets where the minimum margin adds stops
        w = std::min(w, (double)kDenseMaxWidth);
        const double m = (double)margin_for((uint32_t)std::max(w, 2.0));
        if (density * (2.0 * w + m) <= entries * 1.001) break;
        w = (entries / density - m) / 2.0;
    }
    if (w > (double)kDenseMaxWidth) w = (double)kDenseMaxWidth;
    if (w < (double)kDenseMinWidth) {
        if (!force) return false;
        w = (double)kDenseMinWidth;
    }
    w_out = (uint32_t)w;
    m_out = margin_for(w_out);
    if (force) return true;
    const uint64_t bytes = ((uint64_t)max_coord / w_out + 2) * sizeof(Sector);
    if (sector_table_bytes <= (40ull << 20)) return false;  // the sector table is L2-resident: nothing to gain
    if (bytes <= (128ull << 20) && bytes * 5 <= sector_table_bytes * 4) return true;
    return sector_table_bytes > (192ull << 20) && m_out > sector_margin && bytes <= 2 * sector_table_bytes;
}

// Choose up to kMaxClasses contiguous groups of length bins minimising the expected number of
// interval slots find() has to look at per query: sum_c n_c * (maxlen_c + q) / span + kappa per class
// (each class costs two directory lookups).  33 bins, a tiny DP on the host.
void choose_classes(const BuildStats &h, uint64_t span, uint8_t class_of_bin[33], uint32_t &ncls, uint32_t cls_n[4],
                    uint32_t cls_maxlen[4]) {
    const double kQueryLen = 256.0, kKappa = 12.0;
    int bins[33], nbins = 0;
    for (int b = 0; b < 33; b++)
        if (h.count[b]) bins[nbins++] = b;
    auto group_cost = [&](int i, int j) {  // bins[i..j]
        double cnt = 0, mx = 0;
        for (int k = i; k <= j; k++) {
            cnt += (double)h.count[bins[k]];
            mx = std::max(mx, (double)h.maxlen[bins[k]]);
        }
        return cnt * (mx + kQueryLen) / (double)std::max<uint64_t>(span, 1) + kKappa;
    };
    const int G = (int)kMaxClasses;
    double f[5][34];
    int cut[5][34];
    for (int g = 0; g <= G; g++)
        for (int j = 0; j <= nbins; j++) f[g][j] = 1e300, cut[g][j] = -1;
    f[0][0] = 0;
    for (int g = 1; g <= G; g++)
        for (int j = 1; j <= nbins; j++)
            for (int i = g - 1; i < j; i++) {
                if (f[g - 1][i] >= 1e299) continue;
                const double c = f[g - 1][i] + group_cost(i, j - 1);
                if (c < f[g][j]) f[g][j] = c, cut[g][j] = i;
            }
    int best_g = 1;
    for (int g = 1; g <= G && g <= nbins; g++)
        if (f[g][nbins] < f[best_g][nbins]) best_g = g;
    for (int b = 0; b < 33; b++) class_of_bin[b] = 0;
    for (int c = 0; c < 4; c++) cls_n[c] = 0, cls_maxlen[c] = 0;
    ncls = (uint32_t)best_g;
    int j = nbins;
    for (int g = best_g; g >= 1; g--) {
        const int i = cut[g][j];
        for (int k = i; k < j; k++) {
            class_of_bin[bins[k]] = (uint8_t)(g - 1);
            cls_n[g - 1] += (uint32_t)h.count[bins[k]];
            cls_maxlen[g - 1] = std::max(cls_maxlen[g - 1], h.maxlen[bins[k]]);
        }
        j = i;
    }
    // Adjacent classes that are both scan-heavy (more than ~1000 expected window slots per query) become one: no
    // lane ever walks such a window -- their queries get a warp in find_heavy_kernel, which streams ONE class at
    // 128 entries per step but pays a position merge per 32 entries when several are active (cfg 4: the long
    // intervals fall into two length bins; as two classes the merge was 80 % of that kernel's instructions).
    {
        const double kScanHeavy = 1024.0;
        auto slots = [&](int c) {
            return (double)cls_n[c] * ((double)cls_maxlen[c] + kQueryLen) / (double)std::max<uint64_t>(span, 1);
        };
        for (int c = 0; c + 1 < (int)ncls;) {
            if (slots(c) > kScanHeavy && slots(c + 1) > kScanHeavy) {
                cls_n[c] += cls_n[c + 1];
                cls_maxlen[c] = std::max(cls_maxlen[c], cls_maxlen[c + 1]);
                for (int k = c + 1; k + 1 < (int)ncls; k++) cls_n[k] = cls_n[k + 1], cls_maxlen[k] = cls_maxlen[k + 1];
                cls_n[ncls - 1] = 0, cls_maxlen[ncls - 1] = 0;
                for (int b = 0; b < 33; b++)
                    if (class_of_bin[b] > c) class_of_bin[b]--;
                ncls--;
            } else {
                c++;
            }
        }
    }
    // empty bins between / beyond the groups: attach to the next class above so every bin maps somewhere valid
    int last = 0;
    for (int b = 0; b < 33; b++) {
        if (h.count[b]) last = class_of_bin[b];
        else class_of_bin[b] = (uint8_t)last;
    }
}

void build_find_classes(DeviceIndex &ix, cudaStream_t stream) {
    const uint64_t n = ix.n;
    const uint32_t first_start = ix.stats.first_start, last_start = ix.stats.last_start;
    uint8_t class_of_bin[33];
    choose_classes(ix.stats, (uint64_t)last_start - first_start + 1, class_of_bin, ix.ncls, ix.cls_n, ix.cls_maxlen);
    uint32_t off = 0;
    for (uint32_t c = 0; c < 4; c++) {
        ix.cls_off[c] = off;
        off += ix.cls_n[c];
    }

    ix.cls_starts = pool_alloc<uint32_t>(n, ix.bytes, stream);
    ix.cls_stops = pool_alloc<uint32_t>(n, ix.bytes, stream);
    ix.cls_gidx = pool_alloc<uint32_t>(n, ix.bytes, stream);
    launch_class_partition(ix, class_of_bin, stream);
    // directory: about one bucket per six intervals, at least 2^4 coordinates wide
    uint32_t k = 0;
    const uint64_t target = std::max<uint64_t>(n / 6, 1024);
    while (k < 31 && ((uint64_t)(last_start >> k) + 1) > target) k++;
    ix.cls_shift = k;
    ix.cls_nb = (uint64_t)(last_start >> k) + 1;
    ix.cls_dir = pool_alloc<uint4>(ix.cls_nb + 1, ix.bytes, stream);
    fill_class_dir_kernel<<<grid_for(ix.cls_nb + 1), kThreads, 0, stream>>>(
        ix.cls_starts, ix.cls_off[0], ix.cls_n[0], ix.cls_off[1], ix.cls_n[1], ix.cls_off[2], ix.cls_n[2], ix.cls_off[3],
        ix.cls_n[3], k, ix.cls_nb, ix.cls_dir);
    LAPPER_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void launch_sort_u64(uint64_t *keys, uint64_t *tmp, uint64_t n, cudaStream_t stream) {
    if (n) radix_sort<uint64_t, false>(keys, tmp, nullptr, nullptr, n, 64, stream);
}

void launch_sort_pairs_u64(uint64_t *keys, uint64_t *keys_tmp, uint32_t *vals, uint32_t *vals_tmp, uint64_t n, cudaStream_t stream) {
    if (n) radix_sort<uint64_t, true>(keys, keys_tmp, vals, vals_tmp, n, 64, stream);
}

void launch_sort_u32(uint32_t *keys, uint32_t *tmp, uint64_t n, cudaStream_t stream) {
    if (n > 1) radix_sort<uint32_t, false>(keys, tmp, nullptr, nullptr, n, 32, stream);
}

void free_index(DeviceIndex &ix, cudaStream_t stream) {
    void *arrays[] = {ix.cls_starts, ix.cls_stops, ix.cls_gidx, ix.cls_dir, ix.blkmax32, ix.blkmax1024, ix.starts,
                      ix.stops_by_iv, ix.stops_sorted, ix.perm, ix.prefix_max, ix.buckets, ix.pool, ix.dense};
    for (void *p : arrays) pool_free(p, stream);
    ix = DeviceIndex();
}

static void free_query_tables(DeviceIndex &ix, cudaStream_t stream) {
    void *arrays[] = {ix.cls_starts, ix.cls_stops, ix.cls_gidx, ix.cls_dir, ix.blkmax32, ix.blkmax1024, ix.buckets, ix.pool, ix.dense};
    for (void *p : arrays) pool_free(p, stream);
    ix.cls_starts = ix.cls_stops = ix.cls_gidx = nullptr;
    ix.cls_dir = nullptr;
    ix.blkmax32 = ix.blkmax1024 = nullptr;
    ix.buckets = ix.pool = ix.dense = nullptr;
    ix.shift = 0xFFFFFFFFu;
    ix.nbuckets = ix.npool = ix.dense_nb = 0;
    ix.margin = 0;
    ix.ncls = 0;
    ix.dense_w = ix.dense_m = ix.dense_magic = 0;
    ix.dense_overflow = 0;
    ix.tables_ready = false;
}

// One host sync: the scalars every later decision depends on come back together.
void finish_index_core(DeviceIndex &ix, cudaStream_t stream) {
    const uint64_t n = ix.n;
    ix.max_len = 0;
    ix.has_inverted = false;
    ix.max_coord = 0;
    ix.stats = BuildStats();
    ix.tables_ready = false;
    if (n == 0) {
        ix.tables_ready = true;  // nothing to build
        return;
    }
    Scratch<BuildStats> d_stats(1, stream);
    LAPPER_CUDA_CHECK(cudaMemsetAsync(d_stats.p, 0, sizeof(BuildStats), stream));
    const uint32_t g = (uint32_t)std::min<uint64_t>(div_up(n, kThreads), (uint64_t)sm_count() * 8);
    stats_kernel<<<g, kThreads, 0, stream>>>(ix.starts, ix.stops_by_iv, ix.stops_sorted, ix.perm, n, d_stats.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    if (!ix.prefix_max) ix.prefix_max = pool_alloc<uint32_t>(n, ix.bytes, stream);
    launch_prefix_max(ix.stops_by_iv, ix.prefix_max, n, stream);
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(&ix.stats, d_stats.p, sizeof(BuildStats), cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    ix.max_len = ix.stats.max_len;
    ix.has_inverted = ix.stats.inverted != 0;
    ix.max_coord = std::max(ix.stats.last_start, ix.stats.last_stop);
}

// the packed count table next to an existing sector table (ix.buckets): shape from choose_dense_shape
static void build_packed_table(DeviceIndex &ix, uint32_t flags, cudaStream_t stream) {
    uint32_t dw = 0, dm = 0;
    if (!choose_dense_shape(ix.n, ix.max_coord, flags, (ix.nbuckets + 1 + ix.npool) * sizeof(Sector), ix.margin, ix.query_len_hint, dw, dm))
        return;
    Scratch<unsigned long long> d_over(1, stream);
    LAPPER_CUDA_CHECK(cudaMemsetAsync(d_over.p, 0, 8, stream));
    const PackedParams pp = packed_params(dw, ix.max_coord, dm);
    ix.dense_w = pp.w;
    ix.dense_m = pp.m;
    ix.dense_magic = pp.magic;
    ix.dense_nb = pp.nb;
    ix.dense = pool_alloc<Sector>(ix.dense_nb + 1, ix.bytes, stream);
    fill_dense_kernel<<<grid_for(ix.dense_nb + 1), kThreads, 0, stream>>>(ix.starts, ix.stops_sorted, (uint32_t)ix.n, pp, ix.dense, d_over.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    unsigned long long nover = 0;
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(&nover, d_over.p, 8, cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    ix.dense_overflow = nover;
}

void rebuild_packed_table(DeviceIndex &ix, uint32_t flags, cudaStream_t stream) {
    if (!ix.tables_ready || !ix.buckets || ix.n == 0) return;  // (tables not built yet: build_query_tables will read the hint)
    if (ix.dense) {
        ix.bytes -= (ix.dense_nb + 1) * sizeof(Sector);
        pool_free(ix.dense, stream);
    }
    ix.dense = nullptr;
    ix.dense_nb = 0;
    ix.dense_w = ix.dense_m = ix.dense_magic = 0;
    ix.dense_overflow = 0;
    build_packed_table(ix, flags, stream);
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
}

void build_query_tables(DeviceIndex &ix, uint32_t flags, cudaStream_t stream) {
    const uint64_t n = ix.n;
    if (ix.tables_ready) return;
    if (n == 0) {
        ix.tables_ready = true;
        return;
    }
    try {
        uint32_t shift = 0;
        const bool with_buckets = choose_shift(n, ix.max_coord, flags, shift);
        Scratch<unsigned long long> d_counters(2, stream);  // child sectors reserved, packed sectors overflowed
        if (with_buckets) {
            // the sector table first: the number of child sectors it reserves is the one value the host has to wait
            // for, and the class / blkmax kernels below run while it does
            ix.shift = shift;
            ix.margin = shift >= 2 ? (1u << shift) / 4 : 0;
            ix.nbuckets = (uint64_t)(ix.max_coord >> shift) + 1;
            LAPPER_REQUIRE(ix.nbuckets < (1ull << 31), "bucket table too large");
            ix.buckets = pool_alloc<Sector>(ix.nbuckets + 1, ix.bytes, stream);
            LAPPER_CUDA_CHECK(cudaMemsetAsync(d_counters.p, 0, 16, stream));
            fill_buckets_kernel<<<grid_for(ix.nbuckets + 1), kThreads, 0, stream>>>(
                ix.starts, ix.stops_sorted, (uint32_t)n, shift, ix.margin, ix.nbuckets, ix.buckets, d_counters.p);
            LAPPER_CUDA_CHECK(cudaGetLastError());
        }
        build_find_classes(ix, stream);
        {
            const uint64_t nb32 = div_up(n, 32), nb1024 = div_up(nb32, 32);
            ix.blkmax32 = pool_alloc<uint32_t>(nb32, ix.bytes, stream);
            ix.blkmax1024 = pool_alloc<uint32_t>(nb1024, ix.bytes, stream);
            block_max_kernel<<<(uint32_t)div_up(nb32 * 32, kThreads), kThreads, 0, stream>>>(ix.stops_by_iv, n, ix.blkmax32, nb32);
            block_max_kernel<<<(uint32_t)div_up(nb1024 * 32, kThreads), kThreads, 0, stream>>>(ix.blkmax32, nb32, ix.blkmax1024, nb1024);
            LAPPER_CUDA_CHECK(cudaGetLastError());
        }
        if (with_buckets) {
            unsigned long long npool = 0;
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(&npool, d_counters.p, 8, cudaMemcpyDeviceToHost, stream));
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
            LAPPER_REQUIRE(npool < 0xFFFFFFFFull, "child sector pool too large");
            ix.npool = npool;
            if (npool) {
                ix.pool = pool_alloc<Sector>(npool, ix.bytes, stream);
                fill_children_kernel<<<grid_for(ix.nbuckets + 1), kThreads, 0, stream>>>(
                    ix.starts, ix.stops_sorted, (uint32_t)n, shift, ix.nbuckets, ix.buckets, ix.pool);
                LAPPER_CUDA_CHECK(cudaGetLastError());
            }
            build_packed_table(ix, flags, stream);
        }
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));  // the tables are complete when tables_ready says so
        ix.tables_ready = true;
    } catch (...) {
        free_query_tables(ix, stream);
        throw;
    }
}

void build_index_from_unsorted(DeviceIndex &ix, const uint32_t *d_starts, const uint32_t *d_stops, uint64_t n,
                               uint32_t flags, cudaStream_t stream) {
    LAPPER_REQUIRE(n < 0xFFFFFFF0ull, "lapper_build: n must be < 2^32 - 16 (positions are u32)");
    ix = DeviceIndex();
    ix.n = n;
    if (n == 0) {
        finish_index_core(ix, stream);
        return;
    }
    try {
        ix.starts = pool_alloc<uint32_t>(n, ix.bytes, stream);
        ix.stops_by_iv = pool_alloc<uint32_t>(n, ix.bytes, stream);
        ix.stops_sorted = pool_alloc<uint32_t>(n, ix.bytes, stream);
        ix.perm = pool_alloc<uint32_t>(n, ix.bytes, stream);
        {
            Scratch<uint64_t> keys(n, stream), keys_tmp(n, stream);
            Scratch<uint32_t> pos_tmp(n, stream), stops_tmp(n, stream);
            pack_keys_kernel<<<grid_for(n), kThreads, 0, stream>>>(d_starts, d_stops, n, keys.p, ix.perm);
            LAPPER_CUDA_CHECK(cudaGetLastError());
            // (start, stop) order, stable: 8 passes over the 64-bit key, payload = input position
            radix_sort<uint64_t, true>(keys.p, keys_tmp.p, ix.perm, pos_tmp.p, n, 64, stream);
            unpack_keys_kernel<<<grid_for(n), kThreads, 0, stream>>>(keys.p, n, ix.starts, ix.stops_by_iv);
            LAPPER_CUDA_CHECK(cudaGetLastError());
            // the reference's private `stops`: the stops sorted on their own (src/lib.rs:215)
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(ix.stops_sorted, d_stops, n * 4, cudaMemcpyDeviceToDevice, stream));
            radix_sort<uint32_t, false>(ix.stops_sorted, stops_tmp.p, nullptr, nullptr, n, 32, stream);
        }
        finish_index_core(ix, stream);
        build_query_tables(ix, flags, stream);
    } catch (...) {
        free_index(ix, stream);
        throw;
    }
}

}  // namespace lap
