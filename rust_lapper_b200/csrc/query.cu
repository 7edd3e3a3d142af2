// query.cu -- the batched query path: Lapper::count and Lapper::find on the device.
//
//   count(a, b) = lower_bound(starts, b) - upper_bound(stops_sorted, a)       (src/lib.rs:611-618;
//                 `start + 1` there turns bsearch_seq's "<" into "<=", :614), wrapping like usize.
//   find(a, b)  = ascending i in [lower_bound(starts, a (-) max_len), lower_bound(starts, b))
//                 with stops_by_iv[i] > a                                      (src/lib.rs:629-639, :704-717)
//
// The reference walks each search with ~24 dependent cache misses (bsearch_seq_ref, :448-466).
// Here a query normally costs ONE 32-byte sector (index.cuh) fetched with one LDG.E.256; the sorted
// arrays are touched only for overflowing buckets.  Without a bucket table (tiny or degenerate
// sets) the kernels bisect the sorted arrays under a sampled-key directory held in shared memory.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "query_views.cuh"

namespace lap {

// kernels launched by the query path in this process (lapper_query_kernel_launches(): what bench.py reports as
// `gpu_launches`, counted rather than derived)
std::atomic<uint64_t> g_query_launches{0};
std::atomic<uint64_t> g_sorted_min_queries{4096ull * 1024};
#ifdef LAPPER_DEBUG_CHECKS
__device__ unsigned long long g_debug_violations = 0;
#endif

namespace {


#ifndef LAPPER_MIN_CTAS
#define LAPPER_MIN_CTAS 3
#endif
#ifndef LAPPER_QPT
#define LAPPER_QPT 4
#endif
constexpr int kThreads = 256;
constexpr int kQPT = LAPPER_QPT;                   // queries per thread: eight independent sector gathers in flight
constexpr int kTileQ = kThreads * kQPT;   // queries per CTA tile

// ---- count, bucket path, full tiles.  Persistent CTAs walk tiles of 2048 queries; each thread owns
// eight consecutive queries: two 256-bit loads (starts, stops), eight independent sector gathers
// (LDG.E.256), a second predicated round for the ~15 % of queries whose start falls in another
// bucket than their stop, one 256-bit store.  Queries that meet an overflowing bucket are NOT
// resolved in line (that would serialise a divergent search behind every warp): their index goes to
// a per-warp shared-memory queue, and whenever a warp has 32 of them it resolves them one per lane.
// No CTA barrier anywhere in the loop.
constexpr int kWarps = kThreads / 32;
constexpr int kWQCap = 32 + 32 * kQPT;  // leftovers below one drain + one full iteration of deferrals

constexpr int kFindQPT = 4;
constexpr int kFindTile = 256 * kFindQPT;  // queries per tile of the offsets scan
static_assert(kFindTile == kTileQ, "find's offsets tiles must coincide with the count kernel's tiles");

template <bool kOut64, bool kFind>
__device__ __forceinline__ void count_deferred(const BucketView &v, const FindView &f, uint64_t base,
                                               const uint32_t *entry, uint32_t *__restrict__ o32,
                                               uint64_t *__restrict__ o64, uint64_t *__restrict__ tile_sums) {
    const uint64_t j = base + entry[0];
    if (kFind) {
        // find mode: the number of positions find() yields (BITS count when it is exact, the predicate otherwise)
        const uint32_t c = find_count_one(f, entry[1], entry[2]);
        o32[j] = c;
        if (c) atomicAdd(reinterpret_cast<unsigned long long *>(tile_sums + j / kFindTile), (unsigned long long)c);
    } else {
        uint32_t lb, ub;
        bits_count_slow(v, entry[1], entry[2], lb, ub);
        store_count<kOut64>(o32, o64, j, lb, ub);
    }
}

// at deferral time: pull the child sector an overflowing parent points at into L2, so that the
// drain (which runs a few iterations later) finds it there
__device__ __forceinline__ void prefetch_child(const BucketView &v, const Sector &par, uint32_t key) {
    if (sector_overflow(par) && par.w[2] != kNoChild) {
        const uint32_t low = key & ((1u << v.shift) - 1u);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(v.pool + par.w[2] + (low >> (v.shift - par.w[3]))));
    }
}

// load the kQPT starts and stops this thread owns in tile `tile`
__device__ __forceinline__ void load_queries(const uint32_t *__restrict__ qs, const uint32_t *__restrict__ qe,
                                             uint64_t j0, Sector &qa, Sector &qb) {
    if (kQPT == 8) {
        qa = ld_stream_v8(qs + j0);
        qb = ld_stream_v8(qe + j0);
    } else {
        const uint64_t pol = l2_policy_evict_first();  // streams must not push the table out of L2
        const uint4 ta = ld_v4_hint(qs + j0, pol), tb = ld_v4_hint(qe + j0, pol);
        qa.w[0] = ta.x, qa.w[1] = ta.y, qa.w[2] = ta.z, qa.w[3] = ta.w;
        qb.w[0] = tb.x, qb.w[1] = tb.y, qb.w[2] = tb.z, qb.w[3] = tb.w;
    }
}

// kFind: "find mode" for the offsets pass of find_batch -- the same counts (u32), plus per-tile sums
// accumulated into tile_sums (zeroed by the caller) and queries outside the validity of BITS for find
// (start >= stop) sent to the deferred path, where they are counted by the overlap predicate.
// Two builds of the kernel differ only in registers per thread: 3 CTAs/SM (80 registers: the packed path keeps four
// gathers and their thresholds in flight -- best for order-less batches: 0.663 vs 0.785 ms at cfg 2) and 4 CTAs/SM
// (64 registers: the issue-bound sector-table path of start-sorted batches gains from 32 warps/SM: 0.335 vs 0.371 ms).
// For large batches both are launched back to back and every CTA of both takes the same decision from the same 32
// sampled query groups (`batch_looks_sorted`), so exactly one of them does the work and the other exits at once:
// no host round trip, no extra pass over the queries.
template <bool kOut64, bool kFind, int kCtas>
__global__ void __launch_bounds__(kThreads, kCtas) count_bucket_kernel(BucketView v, const FindView f,
                                                                   const uint32_t *__restrict__ qs,
                                                                   const uint32_t *__restrict__ qe, uint64_t ntiles,
                                                                   uint32_t *__restrict__ o32,
                                                                   uint64_t *__restrict__ o64,
                                                                   uint64_t *__restrict__ tile_sums, const BatchProbe probe) {
    if (!batch_accepted(probe)) return;
    __shared__ uint32_t wqueue[kWarps][kWQCap * 3];  // (relative query index, start, stop) per entry
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t *const wq = wqueue[warp];
    uint32_t wq_n = 0;  // warp-uniform
    const uint32_t width = 1u << v.shift, mask = width - 1u;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t j0 = tile * kTileQ + (uint64_t)threadIdx.x * kQPT;
        if (tile + gridDim.x < ntiles) {  // next tile's queries -> L2 while this one is counted
            asm volatile("prefetch.global.L2 [%0];" ::"l"(qs + j0 + (uint64_t)gridDim.x * kTileQ));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(qe + j0 + (uint64_t)gridDim.x * kTileQ));
        }
        Sector qa, qb;
        load_queries(qs, qe, j0, qa, qb);
        uint32_t bb[kQPT];
        Sector sec[kQPT];
        uint32_t res[kQPT];
        uint32_t defer = 0, second = 0;
        // ---- packed-table path: taken by a warp whose 128 queries are spread out (no sector would be shared
        // between lanes) and mostly short enough for one packed sector each
        bool dense_tile = false;
        if (v.dense != nullptr) {
            // spread-out test on the thread's own kQPT consecutive queries: start-sorted batches have them within a
            // few sectors of each other (those warps share sectors between lanes and stay on the sector-table path)
            if (__any_sync(0xFFFFFFFFu, qb.w[kQPT - 1] - qb.w[0] + v.local_span >= 2u * v.local_span)) {
                uint32_t xs[kQPT], te[kQPT], bad = 0;
#pragma unroll
                for (int k = 0; k < kQPT; k++) bad |= (uint32_t)(!packed_locate(v.pk, qa.w[k], qb.w[k], bb[k], xs[k], te[k])) << k;
                dense_tile = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(bad)) <= 8u * kQPT;  // <= a quarter fall back
                if (dense_tile) {
#pragma unroll
                    for (int k = 0; k < kQPT; k++) {
                        LAPPER_CHECK(bb[k] <= v.pk.nb);
                        sec[k] = ld_sector_stream(v.dense + bb[k]);
                    }
#pragma unroll
                    for (int k = 0; k < kQPT; k++) {
                        res[k] = packed_count(sec[k].w, xs[k], te[k]);
                        bad |= (sec[k].w[7] >> 31) << k;
                    }
                    defer = bad;
                }
            }
        }
        if (!dense_tile) {
#pragma unroll
        for (int k = 0; k < kQPT; k++) {
            bb[k] = bucket_of(v, qb.w[k]);
            sec[k] = ld_sector_cached(v.buckets + bb[k]);
        }
#pragma unroll
        for (int k = 0; k < kQPT; k++) {
            const uint32_t a = qa.w[k], b = qb.w[k];
            // lower_bound(starts, b) = rs + 12 - #{lanes >= b - base}
            int acc = (int)(sec[k].w[0] + kBucketLanes);
            acc = sub_lanes_ge(sec[k], (b & mask) * 0x10001u, acc);
            defer |= (uint32_t)sector_overflow(sec[k]) << k;
            // The same sector also answers the stop side when a lies in [base - margin, base + w).  (A stop
            // beyond the last key is clamped to the sentinel bucket; both formulas stay right for it: the
            // sentinel has rs = re_hi = n, no starts, and only margin stops, all of which are >= any low(b).)
            // (a <= b keeps an inverted query whose start is within the margin of 2^32 from wrapping a_rel into range)
            const uint32_t a_rel = a - (bb[k] << v.shift) + v.margin;
            if (a_rel < width + v.margin && a <= b) {
                int ub = sub_lanes_ge(sec[k], (width + a_rel + 1u) * 0x10001u, (int)sec[k].w[1]);
                if (a == 0xFFFFFFFFu) ub = 0;  // src/lib.rs:614 wrap
                res[k] = (uint32_t)(acc - ub);
            } else {
                res[k] = (uint32_t)acc;
                second |= 1u << k;
            }
        }
        }
        if (kFind) {
#pragma unroll
            for (int k = 0; k < kQPT; k++) defer |= (uint32_t)(!(qa.w[k] < qb.w[k])) << k;
        }
        if (second) {
            // start far below the stop's bucket (long or inverted query): its own bucket's sector
#pragma unroll
            for (int k = 0; k < kQPT; k++)
                if (second & (1u << k)) sec[k] = ld_sector_cached(v.buckets + bucket_of(v, qa.w[k]));
#pragma unroll
            for (int k = 0; k < kQPT; k++)
                if (second & (1u << k)) {
                    const uint32_t a = qa.w[k];
                    int ub = sub_lanes_ge(sec[k], (width + (a & mask) + v.margin + 1u) * 0x10001u, (int)sec[k].w[1]);
                    if (a == 0xFFFFFFFFu) ub = 0;
                    res[k] -= (uint32_t)ub;
                    defer |= (uint32_t)sector_overflow(sec[k]) << k;
                }
        }
        if (kOut64) {
#pragma unroll
            for (int k = 0; k < kQPT; k += 4) {
                // a negative difference only arises from an inverted query or interval; it must widen the
                // way usize wraps, which needs lb and ub separately -> such queries take the deferred path
                Sector o;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    o.w[2 * q] = res[k + q];
                    o.w[2 * q + 1] = 0;
                    defer |= (uint32_t)((int32_t)res[k + q] < 0) << (k + q);
                }
                st_stream_v8(o64 + j0 + k, o);
            }
        } else if (kQPT == 8) {
            Sector o;
#pragma unroll
            for (int k = 0; k < kQPT; k++) o.w[k] = res[k];
            st_stream_v8(o32 + j0, o);
        } else {
            st_v4_hint(o32 + j0, make_uint4(res[0], res[1], res[2], res[3]), l2_policy_evict_first());
        }
        if (kFind) {
            // this warp's contribution to its tile's hit total (deferred queries add theirs when resolved)
            uint64_t sum = 0;
            uint32_t big = 0;
#pragma unroll
            for (int k = 0; k < kQPT; k++) {
                const uint32_t c = (defer & (1u << k)) ? 0u : res[k];
                sum += c;
                big |= c;
            }
            if (!__any_sync(0xFFFFFFFFu, big >> 24)) {
                // 128 counts below 2^24 each: the warp's total fits 32 bits, one REDUX instead of ten shuffles
                sum = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)sum);
            } else {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
            }
            if (lane == 0 && sum) atomicAdd(reinterpret_cast<unsigned long long *>(tile_sums + tile), (unsigned long long)sum);
        }
        // ---- enqueue this iteration's deferred queries (ballot-free: exclusive scan of per-lane counts)
        const uint32_t mine = __popc(defer);
        if (__any_sync(0xFFFFFFFFu, mine != 0)) {
            uint32_t inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= (uint32_t)o) inc += y;
            }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            uint32_t slot = wq_n + inc - mine;
            const uint32_t rel0 = (uint32_t)(j0 - (uint64_t)blockIdx.x * kTileQ);  // incl. the thread's offset
#pragma unroll
            for (int k = 0; k < kQPT; k++)
                if (defer & (1u << k)) {
                    if (dense_tile)  // the fallback starts from the sector table: pull b's sector into L2
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(v.buckets + bucket_of(v, qb.w[k])));
                    else
                        prefetch_child(v, sec[k], (second & (1u << k)) ? qa.w[k] : qb.w[k]);  // sec[k] is the sector that overflowed (or a harmless one)
                    LAPPER_CHECK(slot < (uint32_t)kWQCap);
                    wq[3 * slot] = rel0 + k;
                    wq[3 * slot + 1] = qa.w[k];
                    wq[3 * slot + 2] = qb.w[k];
                    slot++;
                }
            wq_n += total;
            __syncwarp();  // queue visible to the warp; also orders the vector store above before the drain's stores
            while (wq_n >= 32) {
                wq_n -= 32;
                count_deferred<kOut64, kFind>(v, f, (uint64_t)blockIdx.x * kTileQ, wq + 3 * (wq_n + lane), o32, o64, tile_sums);
            }
            __syncwarp();
        }
    }
    if (lane < wq_n) count_deferred<kOut64, kFind>(v, f, (uint64_t)blockIdx.x * kTileQ, wq + 3 * lane, o32, o64, tile_sums);
}

// ---- count, bucket path, any alignment / ragged tail: one query per thread
template <bool kOut64, bool kFind>
__global__ void __launch_bounds__(kThreads) count_bucket_tail_kernel(BucketView v, const FindView f,
                                                                     const uint32_t *__restrict__ qs,
                                                                     const uint32_t *__restrict__ qe, uint64_t first,
                                                                     uint64_t nq, uint32_t *__restrict__ o32,
                                                                     uint64_t *__restrict__ o64,
                                                                     uint64_t *__restrict__ tile_sums) {
    for (uint64_t j = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nq;
         j += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t a = ld_stream_u32(qs + j), b = ld_stream_u32(qe + j);
        if (kFind) {
            const uint32_t c = find_count_one(f, a, b);
            o32[j] = c;
            if (c) atomicAdd(reinterpret_cast<unsigned long long *>(tile_sums + j / kFindTile), (unsigned long long)c);
        } else {
            uint32_t lb, ub;
            bits_count_slow(v, a, b, lb, ub);
            store_count<kOut64>(o32, o64, j, lb, ub);
        }
    }
}

// ---- count, direct path: bisect the sorted arrays under a sampled-key directory in shared memory
// (every 2^dshift-th key; the top levels of both searches never leave the SM).
constexpr uint32_t kDirSlots = 2048;

struct PlainView {
    const uint32_t *S;
    const uint32_t *E;
    uint32_t n;
    uint32_t dshift;  // directory stride = 2^dshift
    uint32_t nsamp;   // ceil(n / stride) <= kDirSlots
};

__device__ __forceinline__ uint32_t dir_count_lt(const uint32_t *dir, uint32_t key) {
    // #directory keys < key; dir is padded with 0xFFFFFFFF up to kDirSlots (a power of two)
    uint32_t pos = 0;
#pragma unroll
    for (uint32_t step = kDirSlots >> 1; step > 0; step >>= 1) pos += (dir[pos + step - 1] < key) ? step : 0u;
    return pos + (dir[pos] < key ? 1u : 0u);
}
__device__ __forceinline__ uint32_t dir_count_le(const uint32_t *dir, uint32_t nsamp, uint32_t key) {
    uint32_t lo = 0, hi = nsamp;  // padding equals 0xFFFFFFFF and must not count when key == 0xFFFFFFFF
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (dir[mid] <= key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t plain_lb(const PlainView &v, const uint32_t *dirS, uint32_t key) {
    const uint32_t j = min(dir_count_lt(dirS, key), v.nsamp);
    const uint32_t lo = j ? (j - 1) << v.dshift : 0u;
    const uint32_t hi = min(j << v.dshift, v.n);
    return lower_bound_u32(v.S, lo, hi, key);
}
__device__ __forceinline__ uint32_t plain_ub(const PlainView &v, const uint32_t *dirE, uint32_t key) {
    const uint32_t j = dir_count_le(dirE, v.nsamp, key);
    const uint32_t lo = j ? (j - 1) << v.dshift : 0u;
    const uint32_t hi = min(j << v.dshift, v.n);
    return upper_bound_u32(v.E, lo, hi, key);
}

template <bool kOut64>
__global__ void __launch_bounds__(kThreads) count_plain_kernel(PlainView v, const uint32_t *__restrict__ qs,
                                                               const uint32_t *__restrict__ qe, uint64_t nq,
                                                               uint32_t *__restrict__ o32,
                                                               uint64_t *__restrict__ o64) {
    __shared__ uint32_t dirS[kDirSlots];
    __shared__ uint32_t dirE[kDirSlots];
    for (uint32_t i = threadIdx.x; i < kDirSlots; i += blockDim.x) {
        const bool ok = i < v.nsamp;
        dirS[i] = ok ? __ldg(v.S + ((uint64_t)i << v.dshift)) : 0xFFFFFFFFu;
        dirE[i] = ok ? __ldg(v.E + ((uint64_t)i << v.dshift)) : 0xFFFFFFFFu;
    }
    __syncthreads();
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nq; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t a = ld_stream_u32(qs + j), b = ld_stream_u32(qe + j);
        const uint32_t lb = v.n ? plain_lb(v, dirS, b) : 0u;
        uint32_t ub = v.n ? plain_ub(v, dirE, a) : 0u;
        if (a == 0xFFFFFFFFu) ub = 0;  // src/lib.rs:614 wrap
        store_count<kOut64>(o32, o64, j, lb, ub);
    }
}


__device__ __forceinline__ uint64_t block_reduce_sum_u64(uint64_t v, uint64_t *smem /* 8 */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) smem[w] = v;
    __syncthreads();
    uint64_t r = 0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (kThreads / 32) ? smem[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xFFFFFFFFu, r, o);
    }
    return r;  // valid in thread 0
}

// pass 1 (general form; the fused form is count_bucket_kernel<.., kFind = true>): per-query hit counts
// (thread t of a tile owns queries t*4..t*4+3) and per-tile sums
__global__ void __launch_bounds__(kThreads) find_count_kernel(FindView f, const uint32_t *__restrict__ qs,
                                                              const uint32_t *__restrict__ qe, uint64_t nq,
                                                              uint32_t *__restrict__ counts,
                                                              uint64_t *__restrict__ tile_sums) {
    __shared__ uint64_t red[8];
    const uint64_t base = (uint64_t)blockIdx.x * kFindTile + (uint64_t)threadIdx.x * kFindQPT;
    uint64_t local = 0;
#pragma unroll
    for (int k = 0; k < kFindQPT; k++) {
        const uint64_t j = base + k;
        if (j < nq) {
            const uint32_t c = find_count_one(f, ld_stream_u32(qs + j), ld_stream_u32(qe + j));
            counts[j] = c;
            local += c;
        }
    }
    const uint64_t s = block_reduce_sum_u64(local, red);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s;
}

// pass 2: tile bases in two levels instead of one serial scan of ~100 k sums: (a) every block of 256 consecutive
// tiles ("super-tile") turns its tile sums into exclusive prefixes within the super-tile and emits its total;
// (b) one block scans the few hundred super-tile totals.  A tile's base is then super_base[tile / 256] + tile_sums[tile].
constexpr int kSuperTiles = 256;
// (tile_base may be tile_sums itself: in place)
__global__ void __launch_bounds__(kSuperTiles) super_sums_kernel(const uint64_t *tile_sums, uint64_t ntiles, uint64_t *tile_base,
                                                                 uint64_t *__restrict__ super_sums) {
    __shared__ uint64_t warp_tot[kSuperTiles / 32];
    const uint64_t t = (uint64_t)blockIdx.x * kSuperTiles + threadIdx.x;
    const uint64_t mine = t < ntiles ? tile_sums[t] : 0ull;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    uint64_t wbase = 0, all = 0;
#pragma unroll
    for (int i = 0; i < kSuperTiles / 32; i++) {
        wbase += (i < w) ? warp_tot[i] : 0;
        all += warp_tot[i];
    }
    if (t < ntiles) tile_base[t] = wbase + inc - mine;  // exclusive within the super-tile
    if (threadIdx.x == 0) super_sums[blockIdx.x] = all;
}

// exclusive scan of the super-tile totals in place (one block; nsuper is a few hundred) + the grand total
__global__ void __launch_bounds__(1024) scan_super_sums_kernel(uint64_t *__restrict__ super_sums, uint64_t nsuper,
                                                               uint64_t *__restrict__ total_out) {
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint64_t i0 = 0; i0 < nsuper; i0 += 1024) {
        const uint64_t i = i0 + threadIdx.x;
        const uint64_t mine = i < nsuper ? super_sums[i] : 0ull;
        uint64_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        uint64_t wbase = 0, all = 0;
        for (int k = 0; k < 32; k++) {
            wbase += (k < w) ? warp_tot[k] : 0;
            all += warp_tot[k];
        }
        const uint64_t carry = carry_s;
        if (i < nsuper) super_sums[i] = carry + wbase + inc - mine;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + all;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}

// pass 3: offsets[j] = tile base + exclusive scan of the tile's counts.  128 threads per tile, eight counts each (32 bytes
// of loads and 64 bytes of stores in flight per thread: the kernel only streams, so bytes in flight are what it needs).
constexpr int kWoThreads = 128;
constexpr int kWoQPT = kFindTile / kWoThreads;
static_assert(kWoQPT == 8, "write_offsets_kernel: eight counts per thread");
__global__ void __launch_bounds__(kWoThreads, 16) write_offsets_kernel(const uint32_t *__restrict__ counts,
                                                                       const uint64_t *__restrict__ tile_sums,
                                                                       const uint64_t *__restrict__ super_sums, uint64_t nq,
                                                                       uint64_t *__restrict__ offsets, bool vec) {
    __shared__ uint64_t warp_tot[kWoThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kFindTile + (uint64_t)threadIdx.x * kWoQPT;
    uint32_t c[kWoQPT];
    const bool whole = base + kWoQPT <= nq;
    if (whole) {  // (counts is the library's own scratch: 16-byte aligned)
        const uint4 c0 = ld_stream_v4(counts + base), c1 = ld_stream_v4(counts + base + 4);
        c[0] = c0.x, c[1] = c0.y, c[2] = c0.z, c[3] = c0.w, c[4] = c1.x, c[5] = c1.y, c[6] = c1.z, c[7] = c1.w;
    } else {
#pragma unroll
        for (int k = 0; k < kWoQPT; k++) c[k] = base + k < nq ? counts[base + k] : 0u;
    }
    const uint64_t tile_base = super_sums[blockIdx.x / kSuperTiles] + tile_sums[blockIdx.x];
    uint64_t local = 0;
#pragma unroll
    for (int k = 0; k < kWoQPT; k++) local += c[k];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    uint64_t wbase = 0;
#pragma unroll
    for (int i = 0; i < kWoThreads / 32; i++) wbase += (i < w) ? warp_tot[i] : 0;
    uint64_t run = tile_base + wbase + inc - local;
    if (vec && whole) {  // 64 contiguous bytes per thread: four 128-bit stores (offsets 16-byte aligned)
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(offsets + base);
#pragma unroll
        for (int k = 0; k < kWoQPT / 2; k++) {
            const uint64_t r0 = run, r1 = r0 + c[2 * k];
            dst[k] = make_ulonglong2(r0, r1);
            run = r1 + c[2 * k + 1];
        }
    } else {
#pragma unroll
        for (int k = 0; k < kWoQPT; k++) {
            if (base + k < nq) offsets[base + k] = run;
            run += c[k];
        }
    }
}

#ifndef LAPPER_EMIT_CAND
#define LAPPER_EMIT_CAND 64
#endif
constexpr int kEmitCand = LAPPER_EMIT_CAND;         // candidates a warp can sweep cooperatively
constexpr uint32_t kEmitRawMax = 512;                // class-window slots a warp filters down to those candidates
#ifndef LAPPER_EMIT_WORDS
#define LAPPER_EMIT_WORDS 1536
#endif
#ifndef LAPPER_EMIT_MIN_CTAS
#define LAPPER_EMIT_MIN_CTAS 4
#endif
constexpr int kEmitWarpWords = LAPPER_EMIT_WORDS;    // u32 words of shared memory per warp (6 KB; 48 KB per CTA)
constexpr int kEmitStage = kEmitWarpWords - 4 * kEmitCand;  // output slots staged per warp

struct ClassCursor {
    uint32_t cur, end, g;
};

__device__ __forceinline__ void cursor_settle(const FindView &f, int c, ClassCursor &k, uint32_t a) {
    // advance to the next interval of the class whose stop is beyond the query start
    const uint32_t *ce = f.ce + f.cls_off[c];
    while (k.cur < k.end && __ldg(ce + k.cur) <= a) k.cur++;
    k.g = k.cur < k.end ? __ldg(f.cg + f.cls_off[c] + k.cur) : 0xFFFFFFFFu;
}

// lower bounds of key[c] among the starts of class c, all classes at once: the directory loads of every class are
// issued together, then the (at most 16-entry) rank ranges are probed four independent loads at a time -- a lane
// that owns a query on its own pays a few memory round trips instead of one per probed entry
__device__ __forceinline__ void class_lb_all(const FindView &f, const uint32_t key[4], uint32_t r[4]) {
    uint32_t r0[4], hi[4];
    {
        uint4 d0[4], d1[4];
        uint32_t B[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            B[c] = min(key[c] >> f.cls_shift, f.cls_nb);
            d0[c] = __ldg(f.dir + B[c]);
            d1[c] = __ldg(f.dir + min(B[c] + 1, f.cls_nb));
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            r0[c] = dir_get(d0[c], c);
            hi[c] = B[c] < f.cls_nb ? dir_get(d1[c], c) : r0[c];
        }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        r[c] = r0[c];
        if (c >= (int)f.ncls) continue;
        const uint32_t *cs = f.cs + f.cls_off[c];
        if (hi[c] - r0[c] > 16u) {
            r[c] = lower_bound_u32(cs, r0[c], hi[c], key[c]);
            continue;
        }
        uint32_t x = r0[c];
        while (x < hi[c]) {
            uint32_t v[4];
#pragma unroll
            for (uint32_t t = 0; t < 4; t++) v[t] = x + t < hi[c] ? __ldg(cs + x + t) : 0xFFFFFFFFu;
            uint32_t cnt = 0;
#pragma unroll
            for (uint32_t t = 0; t < 4; t++) cnt += (x + t < hi[c]) && (v[t] < key[c]);
            x += cnt;
            if (cnt < 4) break;
        }
        r[c] = x;
    }
}

// per-lane emission, few hits (the common case of a query that owns its lane: random-order batches): the class
// windows are scanned four independent loads at a time, the hits of each class appended (ascending within the
// class), and the <= 32 positions then merged by an insertion sort in place
constexpr uint32_t kEmitSmall = 32;
__device__ __forceinline__ void emit_one_query_small(const FindView &f, uint32_t a, const uint32_t lo[4], const uint32_t hi[4],
                                                     uint32_t *out, uint32_t room) {
    uint32_t k = 0, first_run = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        if (c >= (int)f.ncls) continue;
        const uint32_t *ce = f.ce + f.cls_off[c], *cg = f.cg + f.cls_off[c];
        for (uint32_t i = lo[c]; i < hi[c]; i += 4) {
            uint32_t e[4];
#pragma unroll
            for (uint32_t t = 0; t < 4; t++) e[t] = i + t < hi[c] ? __ldg(ce + i + t) : 0u;
#pragma unroll
            for (uint32_t t = 0; t < 4; t++)
                if (i + t < hi[c] && e[t] > a) {
                    LAPPER_CHECK(k < room);
                    if (k < room) out[k++] = __ldg(cg + i + t);
                }
        }
        if (c == 0) first_run = k;
    }
    for (uint32_t j = first_run; j < k; j++) {  // runs are ascending: only elements of later runs ever move
        const uint32_t v = out[j];
        uint32_t q = j;
        while (q > 0 && out[q - 1] > v) {
            out[q] = out[q - 1];
            q--;
        }
        out[q] = v;
    }
}

// A lane must not walk long class windows on its own (cfg 4: a query near the end of the universe has few hits,
// but its window in the long class is every long interval that starts before it -- 100 k dependent loads): such a
// query goes on the heavy list whatever its hit count, and find_heavy_kernel gives it a warp.
constexpr uint32_t kLaneWindowMax = 1024;
struct HeavyList {
    uint32_t *list;
    unsigned int *count;      // entries in the list (queries with >= kHeavyHits hits + admitted long-window queries)
    unsigned int *admitted;   // long-window queries admitted so far
    uint32_t max_admitted;    // the list has room for total / kHeavyHits + 1 + max_admitted entries
};

// per-lane emission (any query order) of query j: merge of the class hit streams by global position
__device__ __forceinline__ void emit_one_query(const FindView &f, uint32_t a, uint32_t b, uint32_t *out, uint64_t room,
                                               const HeavyList &heavy, uint64_t j) {
    uint32_t lo[4], hi[4];
    {
        uint32_t key_lo[4], key_hi[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            key_lo[c] = a >= f.cls_maxlen[c] ? a - f.cls_maxlen[c] : 0u;
            key_hi[c] = b;
        }
        class_lb_all(f, key_hi, hi);
        class_lb_all(f, key_lo, lo);
    }
    uint32_t wsum = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        if (c >= (int)f.ncls || lo[c] > hi[c]) lo[c] = hi[c];
        wsum += hi[c] - lo[c];
    }
    if (wsum > kLaneWindowMax && atomicAdd(heavy.admitted, 1u) < heavy.max_admitted) {
        heavy.list[atomicAdd(heavy.count, 1u)] = (uint32_t)j;
        return;  // (no room on the list: walk the windows here after all)
    }
    if (f.ncls > 1 && room <= kEmitSmall) {
        emit_one_query_small(f, a, lo, hi, out, (uint32_t)room);
        return;
    }
    if (f.ncls == 1) {  // single class: its positions are the global positions
        for (uint32_t i = lo[0]; i < hi[0] && room; i++)
            if (__ldg(f.ce + i) > a) {
                *out++ = i;
                room--;
            }
        return;
    }
    ClassCursor k[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        k[c].cur = lo[c], k[c].end = hi[c];
        k[c].g = 0xFFFFFFFFu;
        if (c < (int)f.ncls) cursor_settle(f, c, k[c], a);
    }
    while (room) {
        // smallest global position among the class heads
        const uint32_t m01 = min(k[0].g, k[1].g), m23 = min(k[2].g, k[3].g);
        const uint32_t m = min(m01, m23);
        if (m == 0xFFFFFFFFu) break;
        *out++ = m;
        room--;
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (k[c].g == m) {
                k[c].cur++;
                cursor_settle(f, c, k[c], a);
            }
    }
}

// ---- emit.  One query per lane; a warp's 32 consecutive queries own one contiguous slice of the CSR
// `indices` array, so hits are staged in a per-warp shared-memory slice at (offset - warp base) and
// copied out with fully coalesced stores.
// Fast path (start-sorted or otherwise clustered queries): the 32 queries share nearly all their
// candidates, so the warp gathers the UNION of the class windows once -- at most kEmitCand
// intervals --, ranks them by global position (their merged order), and every lane then sweeps that one
// sorted list with broadcast shared-memory reads, appending the candidates that overlap its own query.
// Slow path (union too large: random-order queries, very long intervals): each lane merges its own
// class streams.  Either way the order is the reference iterator's (ascending position).
#ifdef LAPPER_EMIT_STATS
__device__ unsigned long long g_emit_stats[8];  // chunks, fast, staged-slow, unstaged, sum T raw, sum T kept, sum wtotal, empty
#define EMIT_STAT(i, v) do { if (lane == 0) atomicAdd(&g_emit_stats[i], (unsigned long long)(v)); } while (0)
#else
#define EMIT_STAT(i, v) do { } while (0)
#endif
constexpr uint32_t kHeavyHits = 256;  // a query with at least this many hits gets a warp of its own (find_heavy_kernel)

__device__ __forceinline__ uint32_t smem_lower_bound(const uint32_t *a, uint32_t n, uint32_t key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// Two shapes of the per-warp kernel, both launched, one working (decided on the device from the batch's total, which
// every warp reads from offsets[nq]: no host round trip, like the count kernel's two builds):
//   sparse (< kDenseHitsPerQuery hits per query on average): two queries per lane share one candidate union of at most
//       64 intervals, 6 KB of shared memory per warp, four CTAs per SM -- queries without locality, whose lanes emit on
//       their own, need the occupancy, not the shared memory;
//   dense (cfg 5: 50 M intervals, 43 hits per query, every query brings 2 - 3 new intervals): one query per lane (a
//       second one would add more new candidates than it shares), up to 256 candidates and 2048 staged positions per
//       warp, 12 KB per warp, two CTAs per SM.
template <int kQ_, int kCand_, int kWords_, int kMinCtas_, bool kDense_>
struct FillShape {
    static constexpr int kQ = kQ_;                       // queries per lane
    static constexpr int kCand = kCand_;                 // candidates a warp can sweep cooperatively
    static constexpr int kWords = kWords_;               // u32 words of shared memory per warp
    static constexpr int kStage = kWords_ - 4 * kCand_;  // output slots staged per warp
    static constexpr int kMinCtas = kMinCtas_;
    static constexpr bool kDense = kDense_;
    static_assert(3 * kCand_ <= kWords_ - 4 * kCand_, "the raw candidate list lives in the head of the stage");
    static_assert((kWords_ - 4 * kCand_) % 4 == 0, "the sorted list is 16-byte aligned");
};
using FillSparse = FillShape<2, kEmitCand, kEmitWarpWords, LAPPER_EMIT_MIN_CTAS, false>;
using FillDense = FillShape<1, 256, 3072, 2, true>;
constexpr uint64_t kDenseHitsPerQuery = 20;
static_assert(FillSparse::kStage == kEmitStage, "");

template <typename Shape>
__global__ void __launch_bounds__(kThreads, Shape::kMinCtas) find_fill_kernel(FindView f, const uint32_t *__restrict__ qs,
                                                             const uint32_t *__restrict__ qe, uint64_t nq,
                                                             const uint64_t *__restrict__ offsets,
                                                             uint32_t *__restrict__ indices,
                                                             uint32_t *__restrict__ heavy_list,
                                                             unsigned int *__restrict__ heavy_count,
                                                             const uint32_t *__restrict__ tile_list,
                                                             const unsigned int *__restrict__ tile_count,
                                                             unsigned int *__restrict__ heavy_admitted, uint32_t heavy_max_admitted) {
    constexpr int kEmitQ = Shape::kQ, kEmitCand = Shape::kCand, kEmitStage = Shape::kStage, kEmitWarpWords = Shape::kWords;
    if ((offsets[nq] / kDenseHitsPerQuery >= nq) != Shape::kDense) return;  // the other shape's batch (uniform over the grid)
    const HeavyList heavy = {heavy_list, heavy_count, heavy_admitted, heavy_max_admitted};
    extern __shared__ __align__(16) uint32_t smem_all[];
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t *const stage = smem_all + (threadIdx.x >> 5) * kEmitWarpWords;  // kEmitStage output slots ...
    uint4 *const sorted = reinterpret_cast<uint4 *>(stage + kEmitStage);  // ... then kEmitCand x (start, stop, position, -)
    const uint64_t nwarps = (uint64_t)gridDim.x * (kThreads / 32);
    // with a tile list (the tiles find_fill_tile_kernel left over) the chunks are the 16 (32) chunks of each listed tile
    constexpr uint32_t kChunksPerTile = kFindTile / (32 * kEmitQ);
    const uint64_t nchunks = tile_list ? (uint64_t)*tile_count * kChunksPerTile : div_up(nq, 32 * kEmitQ);
    // A chunk's inputs: the offsets of its queries (and the one after each), the queries, the offset after the chunk.
    // The dense shape loads them one chunk ahead (the warp's next chunk, while this one is swept): two of the five
    // dependent memory round trips of a chunk (offsets -> queries -> directory -> class starts -> candidates) leave the
    // critical path.  The sparse shape has no registers to spare for that (64 at four CTAs per SM) and loads the
    // queries only where there are hits.
    struct ChunkIn {
        uint64_t jbase, o0[kEmitQ], o1[kEmitQ], wend;
        uint32_t a[kEmitQ], b[kEmitQ];
    };
    const auto fetch = [&](uint64_t ci, ChunkIn &in) {
        const uint64_t chunk = tile_list ? (uint64_t)tile_list[ci / kChunksPerTile] * kChunksPerTile + ci % kChunksPerTile : ci;
        // lane's q-th query is chunk*32*kEmitQ + q*32 + lane: consecutive lanes hold consecutive queries
        in.jbase = chunk * (32 * kEmitQ);
        if (in.jbase >= nq) return;
        const uint64_t jend = min(in.jbase + 32 * kEmitQ, nq);  // one past the chunk's last query
#pragma unroll
        for (int q = 0; q < kEmitQ; q++) {
            const uint64_t j = in.jbase + q * 32 + lane;
            const bool live = j < nq;
            in.o0[q] = live ? offsets[j] : 0;
            if (Shape::kDense) {
                in.o1[q] = live ? offsets[j + 1] : 0;
                in.a[q] = live ? ld_stream_u32(qs + j) : 0xFFFFFFFFu;
                in.b[q] = live ? ld_stream_u32(qe + j) : 0u;
            } else {
                // offsets[j + 1] is the next lane's o0, except for lane 31 / the last query
                in.o1[q] = __shfl_down_sync(0xFFFFFFFFu, in.o0[q], 1);
                if (live && (lane == 31 || j + 1 == nq)) in.o1[q] = offsets[j + 1];
                if (!live) in.o0[q] = in.o1[q] = 0;
            }
        }
        in.wend = offsets[jend];  // same address for the whole warp: one broadcast load
    };
    const uint64_t ci0 = (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    ChunkIn cur, nxt;
    if (Shape::kDense && ci0 < nchunks) fetch(ci0, nxt);
    for (uint64_t ci = ci0; ci < nchunks; ci += nwarps) {
        if (Shape::kDense) {
            cur = nxt;
            if (ci + nwarps < nchunks) fetch(ci + nwarps, nxt);
        } else {
            fetch(ci, cur);
        }
        const uint64_t jbase = cur.jbase;
        if (jbase >= nq) continue;
        uint64_t o0[kEmitQ], o1[kEmitQ];
        uint32_t a[kEmitQ], b[kEmitQ];
        bool mine[kEmitQ];
#pragma unroll
        for (int q = 0; q < kEmitQ; q++) {
            const uint64_t j = jbase + q * 32 + lane;
            const bool live = j < nq;
            o0[q] = cur.o0[q], o1[q] = cur.o1[q];
            mine[q] = live && o1[q] > o0[q];
            if (mine[q] && o1[q] - o0[q] >= kHeavyHits) {
                // too many hits for one lane: hand the query to find_heavy_kernel (a warp per query)
                heavy_list[atomicAdd(heavy_count, 1u)] = (uint32_t)j;
                mine[q] = false;
            }
            a[q] = 0xFFFFFFFFu, b[q] = 0;
            if (mine[q]) {
                a[q] = Shape::kDense ? cur.a[q] : ld_stream_u32(qs + j);
                b[q] = Shape::kDense ? cur.b[q] : ld_stream_u32(qe + j);
            }
        }
        const uint64_t wbase = __shfl_sync(0xFFFFFFFFu, o0[0], 0);
        const uint64_t wend = cur.wend;
        const uint64_t wtotal = wend - wbase;
        EMIT_STAT(0, 1);
        if (wtotal == 0) { EMIT_STAT(7, 1); continue; }
        EMIT_STAT(6, wtotal);
        // The stage mirrors the 16-byte alignment of the warp's slice of `indices` (pad words in front), so the
        // copy-out below moves whole 128-bit groups.
        uint32_t *const wdst = indices + wbase;
        const uint32_t pad = (uint32_t)(reinterpret_cast<uintptr_t>(wdst) >> 2) & 3u;
        const bool staged = wtotal + pad <= kEmitStage;
        // ---- union of the class windows over the queries that have hits
        uint32_t amin = a[0], bmax = b[0];
#pragma unroll
        for (int q = 1; q < kEmitQ; q++) amin = min(amin, a[q]), bmax = max(bmax, b[q]);
        amin = __reduce_min_sync(0xFFFFFFFFu, amin);
        bmax = __reduce_max_sync(0xFFFFFFFFu, bmax);
        // Dense shape, first choice: the candidates straight from the single sorted vector -- the reference's own window
        // [lower_bound(amin - max_len), lower_bound(bmax)) (src/lib.rs:632-635) -- when that window is short (sets
        // without long intervals: cfg 5).  It is in position order already: no class directory, no merge of class
        // lists, two loads per entry instead of three.
        uint32_t T = 0, Tk = 0;  // candidates found / kept (warp-uniform)
        bool fits = false, single = false;
        if (Shape::kDense && staged && amin != 0xFFFFFFFFu) {
            uint32_t rr = 0;
            if (lane == 0) rr = rank_lb_starts(f, bmax);
            if (lane == 1) rr = rank_lb_starts(f, amin >= f.max_len ? amin - f.max_len : 0u);
            const uint32_t whi = __shfl_sync(0xFFFFFFFFu, rr, 0), wlo = __shfl_sync(0xFFFFFFFFu, rr, 1);
            const uint32_t n1 = whi > wlo ? whi - wlo : 0u;
            if (n1 <= kEmitRawMax) {
                single = true;
                T = n1;
                for (uint32_t i0 = 0; i0 < n1; i0 += 128) {
                    uint32_t cs[4], ce[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t i = i0 + 32 * u + lane;
                        cs[u] = ce[u] = 0;
                        if (i < n1) {
                            ce[u] = __ldg(f.Eiv + wlo + i);
                            cs[u] = __ldg(f.S + wlo + i);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t i = i0 + 32 * u + lane;
                        const bool keep = i < n1 && ce[u] > amin;
                        const uint32_t km = __ballot_sync(0xFFFFFFFFu, keep);
                        const uint32_t pos = Tk + __popc(km & ((1u << lane) - 1u));
                        if (keep && pos < (uint32_t)kEmitCand) sorted[pos] = make_uint4(cs[u], ce[u], wlo + i, 0u);
                        Tk += __popc(km);
                    }
                }
                fits = Tk <= (uint32_t)kEmitCand;
                if (fits) T = Tk;
                __syncwarp();
            }
        }
        if (!single) {
            // eight class bounds (4 classes x {window start, window end}), four lanes each: the directory gives a
            // rank range of at most 16 class starts, which the four lanes probe with independent loads (one memory
            // round trip instead of a serial walk)
            uint32_t r = 0;
            {
                const int c = (int)((lane >> 2) & 3u);
                const uint32_t sub = lane & 3u;
                const bool on = c < (int)f.ncls;
                const uint32_t key = (lane >> 4) ? bmax : (amin >= f.cls_maxlen[c] ? amin - f.cls_maxlen[c] : 0u);
                const uint32_t *cs = f.cs + f.cls_off[c];
                uint32_t r0 = 0, hi = 0, cnt = 0;
                if (on) {
                    const uint32_t B = min(key >> f.cls_shift, f.cls_nb);
                    r0 = dir_get(__ldg(f.dir + B), c);
                    hi = B < f.cls_nb ? dir_get(__ldg(f.dir + B + 1), c) : r0;
                    if (hi - r0 <= 16u) {
                        uint32_t v[4];
    #pragma unroll
                        for (uint32_t t = 0; t < 4; t++) {
                            const uint32_t i = r0 + sub * 4 + t;
                            v[t] = i < hi ? __ldg(cs + i) : 0xFFFFFFFFu;
                        }
    #pragma unroll
                        for (uint32_t t = 0; t < 4; t++) cnt += (r0 + sub * 4 + t < hi) && (v[t] < key);
                    }
                }
                cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 1);
                cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 2);
                r = r0 + cnt;
                if (on && hi - r0 > 16u) r = lower_bound_u32(cs, r0, hi, key);
            }
            uint32_t lo[4], cnt[4];
            T = 0;
    #pragma unroll
            for (int c = 0; c < 4; c++) {
                lo[c] = __shfl_sync(0xFFFFFFFFu, r, 4 * c);
                const uint32_t hi = __shfl_sync(0xFFFFFFFFu, r, 16 + 4 * c);
                cnt[c] = hi > lo[c] ? hi - lo[c] : 0u;
                T += cnt[c];
            }
            // raw candidates (class-major) into the head of the stage area, which is still unused -- but only those
            // that can overlap some query of the chunk at all: a class window reaches maxlen_c below the smallest query
            // start, and most of the intervals starting down there end before it (stop <= amin)
            uint32_t *const raw = stage;
            Tk = 0;
            uint32_t kc[4] = {0, 0, 0, 0};  // ... per class (warp-uniform)
            fits = staged && T <= kEmitRawMax;
            if (fits) {
                // (dense shape: the loads of four rounds are issued together -- one memory round trip instead of four)
                constexpr int kU = Shape::kDense ? 4 : 1;
                for (uint32_t i0 = 0; i0 < T; i0 += 32 * kU) {
                    uint32_t g[kU], cs[kU], ce[kU], c[kU];
    #pragma unroll
                    for (int u = 0; u < kU; u++) {
                        const uint32_t i = i0 + 32 * u + lane;
                        g[u] = cs[u] = ce[u] = c[u] = 0;
                        if (i < T) {
                            uint32_t base = 0;  // class of candidate i and the index of that class's first candidate
                            if (i >= cnt[0]) {
                                c[u] = 1, base = cnt[0];
                                if (i >= base + cnt[1]) {
                                    c[u] = 2, base += cnt[1];
                                    if (i >= base + cnt[2]) c[u] = 3, base += cnt[2];
                                }
                            }
                            const uint32_t p = f.cls_off[c[u]] + lo[c[u]] + (i - base);
                            ce[u] = __ldg(f.ce + p);  // three independent loads
                            g[u] = __ldg(f.cg + p);
                            cs[u] = __ldg(f.cs + p);
                        }
                    }
    #pragma unroll
                    for (int u = 0; u < kU; u++) {
                        const bool keep = i0 + 32 * u + lane < T && ce[u] > amin;
                        const uint32_t km = __ballot_sync(0xFFFFFFFFu, keep);
                        const uint32_t pos = Tk + __popc(km & ((1u << lane) - 1u));
                        if (keep && pos < (uint32_t)kEmitCand) {
                            raw[pos] = g[u];
                            raw[kEmitCand + pos] = cs[u];
                            raw[2 * kEmitCand + pos] = ce[u];
                        }
                        Tk += __popc(km);
    #pragma unroll
                        for (uint32_t cc = 0; cc < 4; cc++) kc[cc] += __popc(__ballot_sync(0xFFFFFFFFu, keep && c[u] == cc));
                    }
                }
                fits = Tk <= (uint32_t)kEmitCand;
            }
            EMIT_STAT(4, T);
            if (fits) {
                T = Tk;
                __syncwarp();
                // merged order = rank of the global position among all candidates (positions are distinct): the kept list
                // is class-major and ascending within a class, so a candidate's rank is its index in its own class plus
                // one binary search per other class
                // (kR candidates per lane side by side, every search in the same number of power-of-two steps: the
                // shared-memory loads of a step are independent of one another)
                constexpr int kR = Shape::kDense ? 4 : 2;
                for (uint32_t i0 = 0; i0 < T; i0 += 32 * kR) {
                    uint32_t g[kR], k[kR], c[kR], rank[kR];
    #pragma unroll
                    for (int u = 0; u < kR; u++) {
                        const uint32_t i = min(i0 + 32 * u + lane, T - 1);
                        c[u] = 0, k[u] = i;
                        if (k[u] >= kc[0]) {
                            c[u] = 1, k[u] -= kc[0];
                            if (k[u] >= kc[1]) {
                                c[u] = 2, k[u] -= kc[1];
                                if (k[u] >= kc[2]) c[u] = 3, k[u] -= kc[2];
                            }
                        }
                        g[u] = raw[i];
                        rank[u] = k[u];
                    }
                    uint32_t base = 0;
    #pragma unroll
                    for (uint32_t cc = 0; cc < 4; cc++) {
                        const uint32_t n = kc[cc];
                        if (n) {  // (warp-uniform)
                            uint32_t pos[kR];
    #pragma unroll
                            for (int u = 0; u < kR; u++) pos[u] = 0;
                            for (uint32_t step = 1u << (31 - __clz(n)); step; step >>= 1) {
    #pragma unroll
                                for (int u = 0; u < kR; u++)
                                    if (pos[u] + step <= n && raw[base + pos[u] + step - 1] < g[u]) pos[u] += step;
                            }
    #pragma unroll
                            for (int u = 0; u < kR; u++)
                                if (c[u] != cc) rank[u] += pos[u];
                        }
                        base += n;
                    }
    #pragma unroll
                    for (int u = 0; u < kR; u++) {
                        const uint32_t i = i0 + 32 * u + lane;
                        if (i < T) {
                            LAPPER_CHECK(rank[u] < T);
                            sorted[rank[u]] = make_uint4(raw[kEmitCand + i], raw[2 * kEmitCand + i], g[u], 0u);
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (fits) {
            EMIT_STAT(1, 1);
            EMIT_STAT(5, Tk);
            // every lane sweeps the sorted candidates for its own queries (broadcast reads)
            // (write cursors as shared-memory addresses, predicated store + predicated add in inline PTX: five
            // instructions per candidate and query instead of the seven the compiler's own form spends)
            uint32_t wp[kEmitQ];
#pragma unroll
            for (int q = 0; q < kEmitQ; q++) wp[q] = (uint32_t)__cvta_generic_to_shared(stage + pad + (uint32_t)(o0[q] - wbase));
            for (uint32_t x = 0; x < T; x++) {
                const uint4 c = sorted[x];  // one broadcast LDS.128: (start, stop, position)
#pragma unroll
                for (int q = 0; q < kEmitQ; q++) {
#ifdef LAPPER_DEBUG_CHECKS
                    if (c.x < b[q] && c.y > a[q]) {
                        const uint32_t lo_a = (uint32_t)__cvta_generic_to_shared(stage + pad);
                        LAPPER_CHECK(wp[q] >= lo_a && wp[q] < lo_a + 4u * (uint32_t)wtotal);
                    }
#endif
                    // src/lib.rs:140-142: start_i < q_stop && stop_i > q_start
                    asm volatile(
                        "{ .reg .pred p;\n"
                        "  setp.lt.u32 p, %2, %3;\n"
                        "  setp.gt.and.u32 p, %4, %5, p;\n"
                        "  @p st.shared.u32 [%0], %1;\n"
                        "  @p add.u32 %0, %0, 4; }"
                        : "+r"(wp[q])
                        : "r"(c.z), "r"(c.x), "r"(b[q]), "r"(c.y), "r"(a[q])
                        : "memory");
                }
            }
            __syncwarp();
        } else if (staged) {
            EMIT_STAT(2, 1);
#pragma unroll
            for (int q = 0; q < kEmitQ; q++)
                if (mine[q]) emit_one_query(f, a[q], b[q], stage + pad + (o0[q] - wbase), o1[q] - o0[q], heavy, jbase + q * 32 + lane);
            __syncwarp();
        } else {
            EMIT_STAT(3, 1);
#pragma unroll
            for (int q = 0; q < kEmitQ; q++)
                if (mine[q]) emit_one_query(f, a[q], b[q], indices + o0[q], o1[q] - o0[q], heavy, jbase + q * 32 + lane);
        }
        if (staged) {
            // copy-out: stage word pad + i -> wdst[i]; whole 128-bit groups except at the two ragged ends
            const uint32_t W = pad + (uint32_t)wtotal;
            uint32_t *const dst0 = wdst - pad;  // 16-byte aligned; words below pad are never touched
            for (uint32_t g4 = lane * 4; g4 < W; g4 += 128) {
                const uint4 v = *reinterpret_cast<const uint4 *>(stage + g4);
                if (g4 >= pad && g4 + 4 <= W) {
                    st_stream_v4(dst0 + g4, v);
                } else {
                    if (g4 + 0 >= pad && g4 + 0 < W) dst0[g4 + 0] = v.x;
                    if (g4 + 1 >= pad && g4 + 1 < W) dst0[g4 + 1] = v.y;
                    if (g4 + 2 >= pad && g4 + 2 < W) dst0[g4 + 2] = v.z;
                    if (g4 + 3 >= pad && g4 + 3 < W) dst0[g4 + 3] = v.w;
                }
            }
            __syncwarp();
        }
    }
}

// ---- emit, tile form.  A CTA owns a tile of 1024 consecutive queries (the tile of the offsets scan): the tile's
// offsets are loaded once (coalesced, as 32-bit tile-relative values in shared memory), the candidates of the WHOLE
// tile -- the class windows, filtered to intervals that end after the tile's smallest query start -- are gathered
// once by all threads, merged into position order (each class list is already ordered, so an element's rank is its
// own index plus one binary search per other class) and kept in shared memory.  Each warp then takes 64-query
// groups: it filters the tile list down to its own candidates with a ballot (no global loads, already in order),
// every lane sweeps that list for its two queries, and the group's slice of `indices` leaves through the staged
// 128-bit copy-out.  The serial chain of global-memory round trips (offsets -> queries -> directory -> class starts
// -> candidates) is paid once per 1024 queries instead of once per 64, and small CTAs (4 warps) keep seven tiles in
// flight per SM so one tile's chain hides behind the other tiles' sweeps.  Tiles whose candidate set does not fit
// (queries without locality, very long class windows) are appended to `fallback_tiles` and left to find_fill_kernel.
#ifndef LAPPER_TILE_THREADS
#define LAPPER_TILE_THREADS 128
#endif
#ifndef LAPPER_TILE_CAND
#define LAPPER_TILE_CAND 384
#endif
#ifndef LAPPER_TILE_CTAS
#define LAPPER_TILE_CTAS 7
#endif
#ifdef LAPPER_NO_TILE_GUARD  // test-of-the-test builds only
#define TILE_GUARD(x) true
#else
#define TILE_GUARD(x) (x)
#endif
constexpr int kTT = LAPPER_TILE_THREADS;   // threads per CTA
constexpr int kTW = kTT / 32;              // warps per CTA
constexpr int kTQPT = kFindTile / kTT;     // queries per thread in the tile prologue
constexpr int kTCand = LAPPER_TILE_CAND;   // filtered candidates per tile
constexpr uint32_t kTRawMax = 4096;        // class-window slots a tile is willing to filter
#ifndef LAPPER_TILE_STAGE
#define LAPPER_TILE_STAGE 1024
#endif
constexpr int kTStage = LAPPER_TILE_STAGE;  // staged output words per warp
#ifndef LAPPER_TILE_QL
#define LAPPER_TILE_QL 2
#endif
constexpr int kTQL = LAPPER_TILE_QL;        // queries per lane in a group
constexpr uint32_t kTG = 32u * kTQL;        // queries per group
constexpr int kTWarpWords = kTStage + 4 * kEmitCand;  // + the warp's own candidate list (64 x uint4)
constexpr int kTRelWords = kFindTile + 4;
constexpr int kTCtlWords = 32;
constexpr int kTSmemWords = 4 * kTCand + kTW * kTWarpWords + kTRelWords + kTCtlWords;
static_assert(3 * kTCand <= kTW * kTWarpWords, "the raw candidate list aliases the warps' areas");
static_assert(kTQPT == 4 || kTQPT == 8, "tile prologue: 4 or 8 queries per thread");
static_assert(kTW <= 8, "ctl[16 + warp] and ctl[24 + warp]");

// The fused form (kFromCounts, lapper_find_batch_u32_dev): the tile's offsets do not exist yet -- the kernel derives them
// from the count pass's u32 counts and the scanned tile sums (one block scan per tile, which it needs anyway for its
// tile-relative offsets), writes them out for the caller and the later kernels, and emits nothing at all when the
// grand total exceeds the room in `indices`.  That replaces write_offsets_kernel and its 12 bytes per query.
struct TileCounts {
    const uint32_t *counts;      // per query (the count pass)
    const uint64_t *tile_sums;   // per tile: sum of its counts
    const uint64_t *tile_base;   // per tile: exclusive prefix within its super-tile
    const uint64_t *super_base;  // per super-tile: exclusive prefix
    uint64_t *offsets_out;       // offsets[j] for every j < nq are written here (offsets[nq] by scan_super_sums_kernel)
    uint64_t capacity;           // room in `indices`
};

template <bool kFromCounts>
__global__ void __launch_bounds__(kTT, LAPPER_TILE_CTAS) find_fill_tile_kernel(FindView f, const uint32_t *__restrict__ qs,
                                                                    const uint32_t *__restrict__ qe, uint64_t nq,
                                                                    const uint64_t *offsets, const TileCounts tc,
                                                                    uint32_t *__restrict__ indices,
                                                                    uint32_t *__restrict__ heavy_list,
                                                                    unsigned int *__restrict__ heavy_count,
                                                                    uint32_t *__restrict__ fallback_tiles,
                                                                    unsigned int *__restrict__ fallback_count,
                                                                    unsigned int *__restrict__ heavy_admitted, uint32_t heavy_max_admitted) {
    const HeavyList heavy = {heavy_list, heavy_count, heavy_admitted, heavy_max_admitted};
    extern __shared__ __align__(16) uint32_t tsm[];
    uint4 *const cand = reinterpret_cast<uint4 *>(tsm);      // kTCand x (start, stop, position, -)
    uint32_t *const warp_area = tsm + 4 * kTCand;            // per warp: stage, then its candidate list
    uint32_t *const rel = warp_area + kTW * kTWarpWords;     // tile-relative offsets [0, nqt]
    // ctl: [0..3] class lo, [4..7] class window size, [8..11] kept per class, [12] amin, [13] bmax, [14] tile hits >> 32,
    // [16..] warp counts, [24..] the warps' count totals (fused form)
    uint32_t *const ctl = rel + kTRelWords;
    uint32_t *const rawg = warp_area, *const raws = warp_area + kTCand, *const rawe = warp_area + 2 * kTCand;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t *const stage = warp_area + warp * kTWarpWords;
    uint4 *const wlist = reinterpret_cast<uint4 *>(stage + kTStage);
    const uint64_t ntiles = div_up(nq, kFindTile);
    const bool over_capacity = kFromCounts && tc.offsets_out[nq] > tc.capacity;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t q0 = tile * kFindTile;
        const uint32_t nqt = (uint32_t)min((uint64_t)kFindTile, nq - q0);
        const uint32_t jt = threadIdx.x * kTQPT;
        {   // the next tile's offsets (counts) and queries on their way into L2 while this tile is emitted
            const uint64_t qn = q0 + (uint64_t)gridDim.x * kFindTile + jt;
            if (qn + kTQPT <= nq) {
                if (!kFromCounts)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(offsets + qn));
                if ((threadIdx.x & 3u) == 0) {
                    if (kFromCounts) asm volatile("prefetch.global.L2 [%0];" ::"l"(tc.counts + qn));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(qs + qn));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(qe + qn));
                }
            }
        }
        __syncthreads();  // the previous tile's groups are done with the shared arrays
        if (threadIdx.x < 14) ctl[threadIdx.x] = threadIdx.x == 12 ? 0xFFFFFFFFu : 0u;  // [14] is written below
        // ---- 1. the tile's offsets (thread t: queries t*kTQPT .. and the offset after them), heavy queries,
        //         smallest start / largest stop over the queries this kernel emits
        uint64_t tile_o0;
        uint32_t r[kTQPT + 1], qa[kTQPT], qb[kTQPT];
        const bool full = nqt == (uint32_t)kFindTile;
        if (full) {
#pragma unroll
            for (int k = 0; k < kTQPT / 4; k++) {
                // (normal caching, not evict-first: the groups read their queries again a few microseconds later, from L2)
                const uint4 va = __ldg(reinterpret_cast<const uint4 *>(qs + q0 + jt + 4 * k));
                const uint4 vb = __ldg(reinterpret_cast<const uint4 *>(qe + q0 + jt + 4 * k));
                qa[4 * k] = va.x, qa[4 * k + 1] = va.y, qa[4 * k + 2] = va.z, qa[4 * k + 3] = va.w;
                qb[4 * k] = vb.x, qb[4 * k + 1] = vb.y, qb[4 * k + 2] = vb.z, qb[4 * k + 3] = vb.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kTQPT; k++) {
                const bool live = jt + k < nqt;
                qa[k] = live ? ld_stream_u32(qs + q0 + jt + k) : 0xFFFFFFFFu;
                qb[k] = live ? ld_stream_u32(qe + q0 + jt + k) : 0u;
            }
        }
        uint32_t inc = 0, last = 0;  // kFromCounts: the warp's inclusive scan of the threads' sums, the thread's own sum
        uint64_t tile_total = 0;
        if (kFromCounts) {
            tile_total = tc.tile_sums[tile];
            tile_o0 = tc.super_base[tile / kSuperTiles] + tc.tile_base[tile];
            uint32_t c[kTQPT];
            if (full) {
#pragma unroll
                for (int k = 0; k < kTQPT / 4; k++) {
                    const uint4 v = ld_stream_v4(tc.counts + q0 + jt + 4 * k);
                    c[4 * k] = v.x, c[4 * k + 1] = v.y, c[4 * k + 2] = v.z, c[4 * k + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < kTQPT; k++) c[k] = jt + k < nqt ? ld_stream_u32(tc.counts + q0 + jt + k) : 0u;
            }
            // r[k + 1] for now: the thread's own inclusive prefix
            r[0] = 0;
#pragma unroll
            for (int k = 0; k < kTQPT; k++) r[k + 1] = r[k] + c[k];
            last = r[kTQPT];
            inc = last;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= (uint32_t)o) inc += y;
            }
            if (lane == 31) ctl[24 + warp] = inc;
        } else {
            tile_o0 = offsets[q0];
            if (full) {
                const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(offsets + q0 + jt);
#pragma unroll
                for (int k = 0; k < kTQPT / 2; k++) {
                    const ulonglong2 u = src[k];
                    r[2 * k] = (uint32_t)(u.x - tile_o0), r[2 * k + 1] = (uint32_t)(u.y - tile_o0);
                }
                r[kTQPT] = (uint32_t)(offsets[q0 + jt + kTQPT] - tile_o0);
            } else {
#pragma unroll
                for (int k = 0; k <= kTQPT; k++) r[k] = (uint32_t)(offsets[q0 + min(jt + k, nqt)] - tile_o0);
            }
            if (threadIdx.x == kTT - 1) tile_total = offsets[q0 + nqt] - tile_o0;
        }
        if (kFromCounts) {
            __syncthreads();  // ctl initialised, the warps' totals are there
            uint32_t ex = inc - last;
#pragma unroll
            for (int w = 0; w < kTW; w++) ex += w < (int)warp ? ctl[24 + w] : 0u;
#pragma unroll
            for (int k = 0; k <= kTQPT; k++) r[k] += ex;
            // the offsets leave for the caller (and for find_fill_kernel / find_heavy_kernel, which run after this kernel)
            if ((tile_total >> 32) == 0) {
                if (full) {
                    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(tc.offsets_out + q0 + jt);
#pragma unroll
                    for (int k = 0; k < kTQPT / 2; k++) dst[k] = make_ulonglong2(tile_o0 + r[2 * k], tile_o0 + r[2 * k + 1]);
                } else {
#pragma unroll
                    for (int k = 0; k < kTQPT; k++)
                        if (jt + k < nqt) tc.offsets_out[q0 + jt + k] = tile_o0 + r[k];
                }
            } else if (threadIdx.x == 0) {
                // 2^32 or more hits in one tile (possible only over millions of mutually overlapping intervals): 64-bit
                // prefixes, serially; the tile itself is left to find_fill_kernel below (ctl[14] != 0)
                uint64_t run = tile_o0;
                for (uint32_t k = 0; k < nqt; k++) {
                    tc.offsets_out[q0 + k] = run;
                    run += tc.counts[q0 + k];
                }
            }
        }
        uint32_t hmask = 0, amin = 0xFFFFFFFFu, bmax = 0;
#pragma unroll
        for (int k = 0; k < kTQPT; k++) {
            const uint32_t hits = r[k + 1] - r[k];
            if (hits >= kHeavyHits)
                hmask |= 1u << k;
            else if (hits)
                amin = min(amin, qa[k]), bmax = max(bmax, qb[k]);
        }
#pragma unroll
        for (int k = 0; k < kTQPT / 4; k++)
            *reinterpret_cast<uint4 *>(rel + jt + 4 * k) = make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
        if (threadIdx.x == kTT - 1) {
            rel[kFindTile] = r[kTQPT];
            // tile-relative offsets are 32-bit: a tile with 2^32 or more hits is left to find_fill_kernel (64-bit offsets)
            ctl[14] = (uint32_t)(tile_total >> 32);
        }
        amin = __reduce_min_sync(0xFFFFFFFFu, amin);
        bmax = __reduce_max_sync(0xFFFFFFFFu, bmax);
        if (!kFromCounts) __syncthreads();  // ctl initialised
        if (lane == 0) {
            atomicMin(&ctl[12], amin);
            atomicMax(&ctl[13], bmax);
        }
        __syncthreads();
        if (kFromCounts && over_capacity) continue;  // (uniform; the offsets are complete, nothing is emitted)
        if (rel[nqt] == 0 && ctl[14] == 0) continue;
        amin = ctl[12], bmax = ctl[13];
        // ---- 2. class windows of the tile (warp 0; four lanes per bound, as in find_fill_kernel)
        if (warp == 0) {
            const int c = (int)((lane >> 2) & 3u);
            const uint32_t sub = lane & 3u;
            const bool on = c < (int)f.ncls && amin != 0xFFFFFFFFu;
            const uint32_t key = (lane >> 4) ? bmax : (amin >= f.cls_maxlen[c] ? amin - f.cls_maxlen[c] : 0u);
            const uint32_t *cs = f.cs + f.cls_off[c];
            uint32_t r0 = 0, hi = 0, cnt = 0;
            if (on) {
                const uint32_t B = min(key >> f.cls_shift, f.cls_nb);
                r0 = dir_get(__ldg(f.dir + B), c);
                hi = B < f.cls_nb ? dir_get(__ldg(f.dir + B + 1), c) : r0;
                if (hi - r0 <= 16u) {
                    uint32_t v[4];
#pragma unroll
                    for (uint32_t t = 0; t < 4; t++) {
                        const uint32_t i = r0 + sub * 4 + t;
                        v[t] = i < hi ? __ldg(cs + i) : 0xFFFFFFFFu;
                    }
#pragma unroll
                    for (uint32_t t = 0; t < 4; t++) cnt += (r0 + sub * 4 + t < hi) && (v[t] < key);
                }
            }
            cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 1);
            cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 2);
            uint32_t rk = r0 + cnt;
            if (on && hi - r0 > 16u) rk = lower_bound_u32(cs, r0, hi, key);
            const uint32_t rhi = __shfl_down_sync(0xFFFFFFFFu, rk, 16);
            if (lane < 16 && sub == 0) {
                ctl[c] = rk;
                ctl[4 + c] = (on && rhi > rk) ? rhi - rk : 0u;
            }
        }
        __syncthreads();
        // ---- 3. the windows' intervals that end after amin, compacted in class-major order by all threads
        const uint32_t n0 = ctl[4], n1 = ctl[5], n2 = ctl[6], n3 = ctl[7];
        const uint32_t R = n0 + n1 + n2 + n3;
        bool nice = n0 <= kTRawMax && n1 <= kTRawMax && n2 <= kTRawMax && n3 <= kTRawMax && R <= kTRawMax && TILE_GUARD(ctl[14] == 0);
        uint32_t K = 0;  // kept so far (uniform)
        for (uint32_t i0 = 0; nice && i0 < R; i0 += kTT) {
            const uint32_t i = i0 + threadIdx.x;
            uint32_t g = 0, cs = 0, ce = 0, c = 0;
            if (i < R) {
                uint32_t k = i;
                if (k >= n0) {
                    c = 1, k -= n0;
                    if (k >= n1) {
                        c = 2, k -= n1;
                        if (k >= n2) c = 3, k -= n2;
                    }
                }
                const uint32_t p = f.cls_off[c] + ctl[c] + k;
                ce = __ldg(f.ce + p);
                g = __ldg(f.cg + p);
                cs = __ldg(f.cs + p);
            }
            const bool keep = i < R && ce > amin;
            const uint32_t km = __ballot_sync(0xFFFFFFFFu, keep);
            if (lane == 0) ctl[16 + warp] = __popc(km);
            if (keep) atomicAdd(&ctl[8 + c], 1u);
            __syncthreads();
            uint32_t before = 0, all = 0;
#pragma unroll
            for (int w = 0; w < kTW; w++) {
                const uint32_t x = ctl[16 + w];
                before += w < (int)warp ? x : 0u;
                all += x;
            }
            const uint32_t pos = K + before + __popc(km & ((1u << lane) - 1u));
            LAPPER_CHECK(!keep || pos < (uint32_t)kTCand || K + all > (uint32_t)kTCand);
            if (keep && pos < (uint32_t)kTCand) rawg[pos] = g, raws[pos] = cs, rawe[pos] = ce;
            K += all;
            nice = K <= (uint32_t)kTCand;
            __syncthreads();
        }
        if (!nice) {
            if (threadIdx.x == 0) fallback_tiles[atomicAdd(fallback_count, 1u)] = (uint32_t)tile;
            if (warp == 0) EMIT_STAT(2, 1);
            continue;
        }
        if (warp == 0) {
            EMIT_STAT(0, 1);
            EMIT_STAT(1, K);
            EMIT_STAT(3, R);
        }
#pragma unroll
        for (int k = 0; k < kTQPT; k++)
            if (hmask & (1u << k)) heavy_list[atomicAdd(heavy_count, 1u)] = (uint32_t)(q0 + jt + k);
        // ---- 4. merge the class sublists by position (sublist c = [s_c, s_c + k_c) of the raw list)
        {
            const uint32_t k0 = ctl[8], k1 = ctl[9], k2 = ctl[10], k3 = ctl[11];
            for (uint32_t i = threadIdx.x; i < K; i += kTT) {
                uint32_t c = 0, k = i;
                if (k >= k0) {
                    c = 1, k -= k0;
                    if (k >= k1) {
                        c = 2, k -= k1;
                        if (k >= k2) c = 3, k -= k2;
                    }
                }
                const uint32_t g = rawg[i];
                uint32_t rank = k;
                if (c != 0) rank += smem_lower_bound(rawg, k0, g);
                if (c != 1) rank += smem_lower_bound(rawg + k0, k1, g);
                if (c != 2) rank += smem_lower_bound(rawg + k0 + k1, k2, g);
                if (c != 3) rank += smem_lower_bound(rawg + k0 + k1 + k2, k3, g);
                LAPPER_CHECK(rank < K);
                cand[rank] = make_uint4(raws[i], rawe[i], g, 0u);
            }
        }
        __syncthreads();  // cand complete; the raw list (aliasing the warps' areas) is dead
        // ---- 5. groups of kTG consecutive queries, kTQL per lane
#ifdef LAPPER_TILE_XFROM
        // Where the filter below starts: the round of the tile list in which the warp's previous group kept its first
        // candidate.  Everything before that candidate starts before it (the list is in start order), so the previous
        // group dropped it for ending at or before ITS smallest start, and a group whose smallest start is no smaller
        // drops it again.  (Start-sorted batches: the groups' smallest starts ascend.)
        uint32_t xfrom = 0, prev_gamin = 0;
#endif
        for (uint32_t gq = warp * kTG; gq < nqt; gq += kTW * kTG) {
            const uint32_t gq_end = min(gq + kTG, nqt);
            const uint32_t gbase = rel[gq];
            if (rel[gq_end] == gbase) continue;
            uint32_t ro0[kTQL], ro1[kTQL], a[kTQL], b[kTQL];
            bool mine[kTQL];
#pragma unroll
            for (int q = 0; q < kTQL; q++) {
                const uint32_t jl = min(gq + q * 32 + lane, nqt);  // beyond the tile: rel[nqt] twice = no hits
                ro0[q] = rel[jl];
                ro1[q] = rel[min(jl + 1, nqt)];
                const uint32_t hits = ro1[q] - ro0[q];
                mine[q] = hits != 0 && hits < kHeavyHits;
                a[q] = 0xFFFFFFFFu, b[q] = 0;
                if (mine[q]) {
                    a[q] = __ldg(qs + q0 + jl);
                    b[q] = __ldg(qe + q0 + jl);
                }
            }
            uint32_t gamin = a[0], gbmax = b[0];
#pragma unroll
            for (int q = 1; q < kTQL; q++) gamin = min(gamin, a[q]), gbmax = max(gbmax, b[q]);
            gamin = __reduce_min_sync(0xFFFFFFFFu, gamin);
            gbmax = __reduce_max_sync(0xFFFFFFFFu, gbmax);
            // the group's candidates: tile candidates that start before its largest stop and end after its smallest
            // start, in order (the tile list is ordered by position, i.e. by start: stop at the first round past gbmax)
            uint32_t T = 0;
#ifdef LAPPER_TILE_XFROM
            if (gamin < prev_gamin) xfrom = 0;
            prev_gamin = gamin;
            for (uint32_t x0 = xfrom; x0 < K; x0 += 32) {
#else
            for (uint32_t x0 = 0; x0 < K; x0 += 32) {
#endif
                const uint32_t x = x0 + lane;
                uint4 c = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
                if (x < K) c = cand[x];
                const bool keep = c.x < gbmax && c.y > gamin;
                const uint32_t km = __ballot_sync(0xFFFFFFFFu, keep);
                const uint32_t pos = T + __popc(km & ((1u << lane) - 1u));
                if (keep && pos < (uint32_t)kEmitCand) wlist[pos] = c;
#ifdef LAPPER_TILE_XFROM
                if (T == 0 && km) xfrom = x0;
#endif
                T += __popc(km);
                if (__all_sync(0xFFFFFFFFu, c.x >= gbmax)) break;
            }
            const bool fits = T <= (uint32_t)kEmitCand;
            __syncwarp();
            // The group's slice of `indices` goes through the warp's stage in one pass when it fits (nearly always),
            // else in passes over consecutive sub-ranges of its queries; a query that does not fit on its own is
            // heavy (find_heavy_kernel writes it), so every pass makes progress.
            for (uint32_t sub_lo = gq; sub_lo < gq_end;) {
                const uint32_t sbase = rel[sub_lo];
                uint32_t *const wdst = indices + tile_o0 + sbase;
                const uint32_t pad = (uint32_t)(reinterpret_cast<uintptr_t>(wdst) >> 2) & 3u;
                uint32_t sub_hi = gq_end;
                if (rel[gq_end] - sbase + pad > (uint32_t)kTStage) {
                    // largest sub_hi with rel[sub_hi] - sbase + pad <= kTStage (rel is non-decreasing)
                    sub_hi = sub_lo;
#pragma unroll
                    for (int q = 0; q < kTQL; q++) {
                        const uint32_t idx = sub_lo + 1 + q * 32 + lane;
                        const bool ok = idx <= gq_end && rel[min(idx, nqt)] - sbase + pad <= (uint32_t)kTStage;
                        sub_hi += __popc(__ballot_sync(0xFFFFFFFFu, ok));
                    }
                    EMIT_STAT(7, 1);
                    if (sub_hi == sub_lo) {
                        sub_lo++;
                        continue;
                    }
                }
                const uint32_t wtotal = rel[sub_hi] - sbase;
                if (wtotal) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous slice has left the stage
                    __syncwarp();
                    uint32_t sa[kTQL], sb[kTQL];
                    bool act[kTQL];
#pragma unroll
                    for (int q = 0; q < kTQL; q++) {
                        const uint32_t jl = gq + q * 32 + lane;
                        act[q] = mine[q] && jl >= sub_lo && jl < sub_hi;
                        sa[q] = act[q] ? a[q] : 0xFFFFFFFFu;
                        sb[q] = act[q] ? b[q] : 0u;
                    }
                    if (fits) {
                        EMIT_STAT(4, 1);
                        EMIT_STAT(5, T);
                        // write cursors as shared-memory addresses; per candidate and query: two compares, one
                        // predicated store, one predicated add (inline PTX: the compiler's own form spends two
                        // instructions on the predicated cursor update)
                        uint32_t wp[kTQL];
#pragma unroll
                        for (int q = 0; q < kTQL; q++) wp[q] = (uint32_t)__cvta_generic_to_shared(stage + pad + (ro0[q] - sbase));
                        for (uint32_t x = 0; x < T; x++) {
                            const uint4 c = wlist[x];  // one broadcast LDS.128: (start, stop, position)
#pragma unroll
                            for (int q = 0; q < kTQL; q++) {
#ifdef LAPPER_DEBUG_CHECKS
                                if (c.x < sb[q] && c.y > sa[q]) {
                                    const uint32_t lo_a = (uint32_t)__cvta_generic_to_shared(stage + pad);
                                    LAPPER_CHECK(wp[q] >= lo_a && wp[q] < lo_a + 4u * wtotal);
                                }
#endif
                                // src/lib.rs:140-142: start_i < q_stop && stop_i > q_start
                                asm volatile(
                                    "{ .reg .pred p;\n"
                                    "  setp.lt.u32 p, %2, %3;\n"
                                    "  setp.gt.and.u32 p, %4, %5, p;\n"
                                    "  @p st.shared.u32 [%0], %1;\n"
                                    "  @p add.u32 %0, %0, 4; }"
                                    : "+r"(wp[q])
                                    : "r"(c.z), "r"(c.x), "r"(sb[q]), "r"(c.y), "r"(sa[q])
                                    : "memory");
                            }
                        }
                    } else {
                        EMIT_STAT(6, 1);
#pragma unroll
                        for (int q = 0; q < kTQL; q++)
                            if (act[q])
                                emit_one_query(f, a[q], b[q], stage + pad + (ro0[q] - sbase), ro1[q] - ro0[q], heavy,
                                               q0 + gq + q * 32 + lane);
                    }
                    // copy-out: the whole 16-byte groups of the slice leave as ONE bulk copy shared -> global issued by
                    // lane 0 (cp.async.bulk, the TMA path: no LDS/STG instructions, no LSU wavefronts); the ragged
                    // words at the two ends are stored by lanes 0..7.  The stage is reused only after the copy has
                    // finished reading it (wait at the top of the next pass).
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this lane's STS -> visible to the async proxy
                    __syncwarp();
                    const uint32_t W = pad + wtotal;
                    uint32_t *const dst0 = wdst - pad;  // 16-byte aligned; words below pad are never touched
                    const uint32_t hb = pad ? 4u : 0u, tb = W & ~3u;
                    const bool body = tb > hb;
                    if (lane < 8) {
                        const uint32_t idx = (body && lane >= 4) ? tb + lane - 4 : lane;
                        const bool edge = body ? (lane < 4 ? (pad != 0 && idx >= pad) : idx < W) : (idx >= pad && idx < W);
                        if (edge) dst0[idx] = stage[idx];
                    }
                    if (lane == 0 && body) {
                        const uint32_t src = (uint32_t)__cvta_generic_to_shared(stage + hb);
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst0 + hb), "r"(src),
                                     "r"((tb - hb) * 4u)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                sub_lo = sub_hi;
            }
        }
        // the warp's last slice must have left its stage before the next tile's candidate lists overwrite it
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// ---- heavy queries: one warp per query scans the window [lo, hi) of the single sorted vector -- the
// reference's own scan (src/lib.rs:704-717) -- 32 intervals per step, skipping every block of 32 / 1024
// intervals whose max stop is <= the query start, and ballot-compacts the hits into coalesced stores.
// Positions come out ascending by construction: no merge.  Warps fetch queries from an atomic counter.
// Heavy query, sparse case: the hits are a small fraction of the single sorted vector's window (cfg 4: 1 % of
// the intervals are long and spread over the whole vector, so nearly every block of it has to be looked at for
// about one hit).  The length classes hold exactly those intervals contiguously: the warp scans the class windows
// instead -- a block of 32 entries per class and round -- and merges the class hit streams by position on the fly.
// G = the smallest last-loaded position over the classes that continue beyond their block: everything at or below
// G is complete in registers, so those entries are consumed, and an emitted entry's place in the output is its rank
// among the emitted entries of all classes (own class: ballot prefix; other classes: a 6-shuffle lower bound in
// their sorted block, then a popcount of their emit mask).  The class that defines G consumes its whole block, so
// every round makes progress.
__device__ __forceinline__ void heavy_merge_classes(const FindView &f, uint32_t a, uint32_t b, uint32_t *__restrict__ out,
                                                    uint64_t room, uint32_t lane, uint32_t wlo, uint32_t whi) {
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t p[4], end[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        p[c] = __shfl_sync(0xFFFFFFFFu, wlo, c);
        end[c] = __shfl_sync(0xFFFFFFFFu, whi, c);
        if (c >= (int)f.ncls) p[c] = end[c] = 0;
    }
    uint64_t written = 0;
    for (;;) {
        {   // Streaming step.  The class whose next entry has the smallest position can be emitted on its own as far
            // as its positions stay below every other class's next position -- no merge needed for that stretch.
            // (cfg 4: the hits of the short classes sit at the END of the merged order, next to the query, so the
            // long class streams alone almost to its end although the short classes are still pending.)
            uint32_t head[4];
            int alive = 0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                head[c] = 0xFFFFFFFFu;
                if (p[c] < end[c]) {
                    alive++;
                    head[c] = __ldg(f.cg + f.cls_off[c] + p[c]);  // same address in every lane
                }
            }
            if (alive == 0) break;
            int cs = 0;
#pragma unroll
            for (int c = 1; c < 4; c++)
                if (head[c] < head[cs]) cs = c;
            uint32_t limit = 0xFFFFFFFFu;
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (c != cs) limit = min(limit, head[c]);
            const uint32_t *ce = f.ce + f.cls_off[cs], *cg = f.cg + f.cls_off[cs];
            // worth streaming if at least the next 32 entries of the class lie below the limit
            const uint32_t probe = min(end[cs] - p[cs], 32u);
            if (alive == 1 || __ldg(cg + p[cs] + probe - 1) < limit) {
                // 128 entries per step, loaded speculatively: those below the limit (a prefix: positions ascend) are
                // consumed, and the stream goes on without looking at the other classes until one is not
                for (;;) {
                    const uint32_t step = min(end[cs] - p[cs], 128u);
                    uint32_t e[4], gg[4];
                    bool ok[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t k = 32 * u + lane;
                        ok[u] = k < step;
                        e[u] = ok[u] ? __ldg(ce + p[cs] + k) : 0u;
                        gg[u] = ok[u] ? __ldg(cg + p[cs] + k) : 0xFFFFFFFFu;
                    }
                    uint32_t consumed = 0;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        ok[u] = ok[u] && gg[u] < limit;
                        consumed += __popc(__ballot_sync(0xFFFFFFFFu, ok[u]));
                        const bool h = ok[u] && e[u] > a;
                        const uint32_t hm = __ballot_sync(0xFFFFFFFFu, h);
                        const uint64_t at = written + __popc(hm & lt);
                        LAPPER_CHECK(!h || at < room);
                        if (h && at < room) out[at] = gg[u];
                        written += __popc(hm);
                    }
                    p[cs] += consumed;
                    if (consumed < step || p[cs] >= end[cs]) break;
                }
                continue;
            }
        }
        uint32_t g[4], em[4];
        bool hit[4], valid[4];
        uint32_t G = 0xFFFFFFFFu;
        bool any = false;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            g[c] = 0xFFFFFFFFu, hit[c] = false, valid[c] = false, em[c] = 0;
            if (p[c] < end[c]) {  // warp-uniform
                any = true;
                const uint32_t i = p[c] + lane;
                valid[c] = i < end[c];
                if (valid[c]) {
                    g[c] = __ldg(f.cg + f.cls_off[c] + i);
                    hit[c] = __ldg(f.ce + f.cls_off[c] + i) > a;  // start < b holds inside the class window
                }
                if (p[c] + 32 < end[c]) G = min(G, __shfl_sync(0xFFFFFFFFu, g[c], 31));
            }
        }
        if (!any) break;
        uint32_t total = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (p[c] < end[c]) {
                const bool take = valid[c] && g[c] <= G;
                em[c] = __ballot_sync(0xFFFFFFFFu, take && hit[c]);
                p[c] += __popc(__ballot_sync(0xFFFFFFFFu, take));
                total += __popc(em[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (em[c] == 0) continue;  // warp-uniform
            uint32_t r = __popc(em[c] & lt);
#pragma unroll
            for (int c2 = 0; c2 < 4; c2++) {
                if (c2 == c || em[c2] == 0) continue;  // warp-uniform
                // #entries of class c2's block with position < g[c]: lower bound over the 32 lanes (ascending, padded
                // with 0xFFFFFFFF)
                uint32_t pos = 0;
#pragma unroll
                for (uint32_t step = 16; step >= 1; step >>= 1)
                    if (__shfl_sync(0xFFFFFFFFu, g[c2], pos + step - 1) < g[c]) pos += step;
                if (__shfl_sync(0xFFFFFFFFu, g[c2], pos) < g[c]) pos++;
                r += __popc(em[c2] & (pos >= 32 ? 0xFFFFFFFFu : (1u << pos) - 1u));
            }
            if ((em[c] >> lane) & 1u) {
                LAPPER_CHECK(written + r < room);
                if (written + r < room) out[written + r] = g[c];
            }
        }
        written += total;
    }
}

__global__ void __launch_bounds__(kThreads) find_heavy_kernel(FindView f, const uint32_t *__restrict__ qs,
                                                              const uint32_t *__restrict__ qe,
                                                              const uint64_t *__restrict__ offsets,
                                                              uint32_t *__restrict__ indices,
                                                              const uint32_t *__restrict__ heavy_list,
                                                              const unsigned int *__restrict__ heavy_count,
                                                              unsigned int *__restrict__ next) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt = (1u << lane) - 1u;
    const unsigned int total = *heavy_count;
    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(next, 1u);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= total) return;
        const uint64_t j = heavy_list[item];
        const uint32_t a = __ldg(qs + j), b = __ldg(qe + j);
        const uint64_t o0 = offsets[j], room = offsets[j + 1] - o0;
        // window: starts in [a - max_len, b), tightened by the running max of stops (nothing before the
        // first position whose prefix max exceeds a can end after a)
        uint32_t r = 0;
        if (lane == 0) r = rank_lb_starts(f, b);
        if (lane == 1) r = rank_lb_starts(f, a >= f.max_len ? a - f.max_len : 0u);
        if (lane == 2) r = upper_bound_u32(f.pmax, 0, f.n, a);
        const uint32_t hi = __shfl_sync(0xFFFFFFFFu, r, 0);
        const uint32_t lo = max(__shfl_sync(0xFFFFFFFFu, r, 1), __shfl_sync(0xFFFFFFFFu, r, 2));
        uint32_t *out = indices + o0;
        // class windows (lane c: class c); scanning them and merging pays when they are much shorter than [lo, hi)
        uint32_t wlo = 0, whi = 0;
        if (lane < f.ncls) class_window(f, (int)lane, a, b, wlo, whi);
        uint32_t wsum = whi - wlo;
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) wsum += __shfl_xor_sync(0xFFFFFFFFu, wsum, o);
        wsum = __shfl_sync(0xFFFFFFFFu, wsum, 0);
        if (f.ncls > 0 && (uint64_t)wsum * 8 < (uint64_t)(hi > lo ? hi - lo : 0u)) {
            heavy_merge_classes(f, a, b, out, room, lane, wlo, whi);
            continue;
        }
        uint64_t written = 0;
        if (room * 2 >= (uint64_t)(hi > lo ? hi - lo : 0u)) {
            // dense window (README bakeoff shape: every second interval of the window is a hit): no block can be
            // skipped, so stream it, 128 intervals per step with independent loads
            for (uint32_t i0 = lo; i0 < hi; i0 += 128) {
                uint32_t e[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + 32 * u + lane;
                    e[u] = i < hi ? __ldg(f.Eiv + i) : 0u;  // stops are > a >= 0 for a hit; 0 never is
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + 32 * u + lane;
                    const bool h = i < hi && e[u] > a;
                    const uint32_t hm = __ballot_sync(0xFFFFFFFFu, h);
                    const uint64_t at = written + __popc(hm & lt);
                    if (h && at < room) out[at] = i;
                    written += __popc(hm);
                }
            }
            continue;
        }
        for (uint32_t b2 = lo >> 10; ((uint64_t)b2 << 10) < hi && written < room; b2++) {
            if (__ldg(f.blkmax1024 + b2) <= a) continue;  // warp-uniform
            const uint32_t b1 = b2 * 32 + lane;            // this lane's block of 32 intervals
            const uint64_t base1 = (uint64_t)b1 << 5;
            const bool live = base1 < hi && base1 + 32 > lo && __ldg(f.blkmax32 + b1) > a;
            uint32_t m = __ballot_sync(0xFFFFFFFFu, live);
            while (m) {
                // up to four surviving blocks per step: their loads are independent, so the latency of the
                // scattered 128-byte reads overlaps
                uint64_t idx[4];
                bool hit[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool have = m != 0;
                    const uint32_t l1 = have ? __ffs(m) - 1 : 0u;
                    m &= m - 1;  // 0 stays 0
                    idx[u] = (((uint64_t)b2 * 32 + l1) << 5) + lane;
                    hit[u] = have && idx[u] >= lo && idx[u] < hi;  // start < b holds for every i < hi
                }
                uint32_t e[4];
#pragma unroll
                for (int u = 0; u < 4; u++) e[u] = hit[u] ? __ldg(f.Eiv + idx[u]) : 0u;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool h = hit[u] && e[u] > a;
                    const uint32_t hm = __ballot_sync(0xFFFFFFFFu, h);
                    if (h && written + __popc(hm & lt) < room) out[written + __popc(hm & lt)] = (uint32_t)idx[u];
                    written += __popc(hm);
                }
            }
        }
    }
}

}  // namespace

namespace {

// LAPPER_COUNT_NO_DUAL=1 (experiments / tests): one build of the count kernel only
static bool count_dual_disabled() {
    static const bool off = [] {
        const char *e = std::getenv("LAPPER_COUNT_NO_DUAL");
        return e && atoi(e) != 0;
    }();
    return off;
}

// How a batch over the bucket table is split between the kernels (decided once per call, shared by every launch of it).
struct BucketPlan {
    uint64_t ntiles;   // full tiles of kTileQ queries that the tiled kernels take (0: pointers not 32-byte aligned)
    bool dual;         // both register builds of count_bucket_kernel, selected on the device
    bool sorted;       // count_sorted_kernel on offer for the tiled part
    BatchProbe probe;  // .accept is filled in per launch
};
static_assert(kProbeTile == (uint32_t)kTileQ, "the batch probe samples the count kernel's tiles");

BucketPlan plan_buckets(const BucketView &v, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq, const void *out,
                        bool sorted_possible) {
    BucketPlan p;
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_qs) | reinterpret_cast<uintptr_t>(d_qe) |
                           reinterpret_cast<uintptr_t>(out)) & 31u) == 0;
    p.ntiles = aligned ? nq / kTileQ : 0;
    const bool big = p.ntiles >= 4096;  // enough queries for the probe's 32 samples to mean something
    p.dual = big && v.dense != nullptr && v.local_span != 0 && !count_dual_disabled();
    // (>= 32 tiles: the probe takes one sample per 32nd of the batch)
    // (fewer than 2^31 intervals: the sorted-run kernel sign-extends 32-bit rank differences)
    p.sorted = p.ntiles * kTileQ >= g_sorted_min_queries.load(std::memory_order_relaxed) && sorted_possible && v.n < (1u << 31) &&
               !count_sorted_disabled();
    p.probe.qs = d_qs, p.probe.qe = d_qe, p.probe.ntiles = p.ntiles;
    p.probe.n = v.n, p.probe.shift = v.shift, p.probe.nb = v.nb, p.probe.local_span = v.local_span;
    p.probe.sorted_ok = p.sorted ? 1u : 0u;
    p.probe.accept = kAcceptAll;
    return p;
}
BatchProbe probe_for(const BucketPlan &p, uint32_t accept) {
    BatchProbe b = p.probe;
    b.accept = (p.dual || p.sorted) ? accept : kAcceptAll;
    return b;
}

// count over the bucket table.  tile_sums != nullptr selects find mode (u32 counts + per-tile sums).
void launch_count_buckets(const DeviceIndex &ix, const BucketPlan &plan, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq,
                          uint32_t *d_out32, uint64_t *d_out64, uint64_t *tile_sums, cudaStream_t stream) {
    const BucketView v = make_bucket_view(ix);
    const FindView f = make_find_view(ix);
    const bool out64 = d_out64 != nullptr;
    const uint64_t ntiles = plan.ntiles;
    constexpr uint32_t kRandom = 1u << kBatchRandom, kLocal = 1u << kBatchLocal, kSorted = 1u << kBatchSortedDense;
    if (ntiles) {
        // persistent grids: SMs (148 on B200) x resident CTAs of 256 threads.  Large batches: the 4-CTA build for batches
        // with local queries, the 3-CTA build for order-less ones, count_sorted_kernel for sorted dense ones -- all
        // launched, the batch probe lets exactly one of them work.
        const auto launch = [&](auto ctas, uint32_t accept) {
            constexpr int kCtas = decltype(ctas)::value;
            const BatchProbe pb = probe_for(plan, accept);
            const uint32_t grid = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)sm_count() * kCtas);
            if (tile_sums)
                g_query_launches++, count_bucket_kernel<false, true, kCtas><<<grid, kThreads, 0, stream>>>(v, f, d_qs, d_qe, ntiles, d_out32, nullptr, tile_sums, pb);
            else if (out64)
                g_query_launches++, count_bucket_kernel<true, false, kCtas><<<grid, kThreads, 0, stream>>>(v, f, d_qs, d_qe, ntiles, nullptr, d_out64, nullptr, pb);
            else
                g_query_launches++, count_bucket_kernel<false, false, kCtas><<<grid, kThreads, 0, stream>>>(v, f, d_qs, d_qe, ntiles, d_out32, nullptr, nullptr, pb);
            LAPPER_CUDA_CHECK(cudaGetLastError());
        };
        if (plan.dual) {
            launch(std::integral_constant<int, 4>{}, kLocal);
            launch(std::integral_constant<int, LAPPER_MIN_CTAS>{}, kRandom);
        } else {
            launch(std::integral_constant<int, LAPPER_MIN_CTAS>{}, kRandom | kLocal);
        }
        if (plan.sorted)
            launch_count_sorted(ix, d_qs, d_qe, ntiles * (kTileQ / kProbeRun), d_out32, d_out64, tile_sums, probe_for(plan, kSorted), stream);
    }
    const uint64_t done = ntiles * kTileQ, rest = nq - done;
    if (rest) {
        const uint32_t grid = capped_grid(rest, kThreads, 16);
        if (tile_sums)
            g_query_launches++, count_bucket_tail_kernel<false, true><<<grid, kThreads, 0, stream>>>(v, f, d_qs, d_qe, done, nq, d_out32, nullptr, tile_sums);
        else if (out64)
            g_query_launches++, count_bucket_tail_kernel<true, false><<<grid, kThreads, 0, stream>>>(v, f, d_qs, d_qe, done, nq, nullptr, d_out64, nullptr);
        else
            g_query_launches++, count_bucket_tail_kernel<false, false><<<grid, kThreads, 0, stream>>>(v, f, d_qs, d_qe, done, nq, d_out32, nullptr, nullptr);
        LAPPER_CUDA_CHECK(cudaGetLastError());
    }
}

}  // namespace

void launch_count(const DeviceIndex &ix, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq, uint32_t *d_out32,
                  uint64_t *d_out64, cudaStream_t stream) {
    if (nq == 0) return;
    constexpr uint64_t kMaxPerLaunch = 1ull << 31;  // deferred-queue entries are 32-bit relative indices
    if (nq > kMaxPerLaunch) {
        for (uint64_t done = 0; done < nq; done += kMaxPerLaunch) {
            const uint64_t m = std::min(kMaxPerLaunch, nq - done);
            launch_count(ix, d_qs + done, d_qe + done, m, d_out32 ? d_out32 + done : nullptr,
                         d_out64 ? d_out64 + done : nullptr, stream);
        }
        return;
    }
    const bool out64 = d_out64 != nullptr;
    if (ix.buckets) {
        const void *out = out64 ? (const void *)d_out64 : (const void *)d_out32;
        const BucketPlan plan = plan_buckets(make_bucket_view(ix), d_qs, d_qe, nq, out, true);
        launch_count_buckets(ix, plan, d_qs, d_qe, nq, d_out32, d_out64, nullptr, stream);
    } else {
        PlainView v;
        v.S = ix.starts;
        v.E = ix.stops_sorted;
        v.n = (uint32_t)ix.n;
        v.dshift = 0;
        while (div_up(ix.n, 1ull << v.dshift) > kDirSlots) v.dshift++;
        v.nsamp = (uint32_t)div_up(ix.n, 1ull << v.dshift);
        const uint32_t grid = capped_grid(nq, kThreads, 8);
        if (out64)
            g_query_launches++, count_plain_kernel<true><<<grid, kThreads, 0, stream>>>(v, d_qs, d_qe, nq, nullptr, d_out64);
        else
            g_query_launches++, count_plain_kernel<false><<<grid, kThreads, 0, stream>>>(v, d_qs, d_qe, nq, d_out32, nullptr);
    }
    LAPPER_CUDA_CHECK(cudaGetLastError());
}

#ifdef LAPPER_EMIT_STATS
extern "C" void lapper_debug_emit_stats(unsigned long long *out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_emit_stats, sizeof(g_emit_stats));
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cudaMemcpyToSymbol(g_emit_stats, z, sizeof(z));
    }
}
#endif

long long debug_violations() {
#ifdef LAPPER_DEBUG_CHECKS
    unsigned long long v = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&v, g_debug_violations, sizeof(v));
    return (long long)v;
#else
    return -1;
#endif
}

uint64_t launch_find_offsets(const DeviceIndex &ix, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq,
                             uint64_t *d_offsets, cudaStream_t stream) {
    if (nq == 0) {
        LAPPER_CUDA_CHECK(cudaMemsetAsync(d_offsets, 0, sizeof(uint64_t), stream));
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
        return 0;
    }
    const FindView f = make_find_view(ix);
    const uint64_t ntiles = div_up(nq, kFindTile);
    LAPPER_REQUIRE(ntiles < 0x7FFFFFFFull, "find_batch: too many queries in one call");
    Scratch<uint32_t> counts(nq, stream);
    Scratch<uint64_t> tile_sums(ntiles, stream);
    if (ix.n == 0) {
        LAPPER_CUDA_CHECK(cudaMemsetAsync(d_offsets, 0, (nq + 1) * sizeof(uint64_t), stream));
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
        return 0;
    }
    if (!ix.has_inverted && ix.buckets && nq < (1ull << 31)) {
        // every interval has start <= stop: the BITS count IS the number of hits of a non-empty query, so
        // the count kernel runs in find mode (it also accumulates the tile sums and recounts start >= stop
        // queries by the predicate)
        const BucketPlan plan = plan_buckets(f.bv, d_qs, d_qe, nq, counts.p, true);
        LAPPER_CUDA_CHECK(cudaMemsetAsync(tile_sums.p, 0, ntiles * sizeof(uint64_t), stream));
        launch_count_buckets(ix, plan, d_qs, d_qe, nq, counts.p, nullptr, tile_sums.p, stream);
    } else {
        g_query_launches++, find_count_kernel<<<(uint32_t)ntiles, kThreads, 0, stream>>>(f, d_qs, d_qe, nq, counts.p, tile_sums.p);
    }
    LAPPER_CUDA_CHECK(cudaGetLastError());
    const uint64_t nsuper = div_up(ntiles, kSuperTiles);
    Scratch<uint64_t> super_sums(nsuper, stream);
    g_query_launches++, super_sums_kernel<<<(uint32_t)nsuper, kSuperTiles, 0, stream>>>(tile_sums.p, ntiles, tile_sums.p, super_sums.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    g_query_launches++, scan_super_sums_kernel<<<1, 1024, 0, stream>>>(super_sums.p, nsuper, d_offsets + nq);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    g_query_launches++, write_offsets_kernel<<<(uint32_t)ntiles, kWoThreads, 0, stream>>>(
        counts.p, tile_sums.p, super_sums.p, nq, d_offsets, (reinterpret_cast<uintptr_t>(d_offsets) & 15u) == 0);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    uint64_t total = 0;
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(&total, d_offsets + nq, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    return total;
}

// LAPPER_FIND_NO_TILE=1 (experiments / tests): per-warp emit kernel only
static bool find_tile_disabled() {
    static const bool off = [] {
        const char *e = std::getenv("LAPPER_FIND_NO_TILE");
        return e && atoi(e) != 0;
    }();
    return off;
}

__global__ void rebase_offsets_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, uint64_t count, uint64_t base) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = in[i] + base;
}

void launch_rebase_offsets(const uint64_t *in, uint64_t *out, uint64_t count, uint64_t base, cudaStream_t stream) {
    if (count == 0) return;
    rebase_offsets_kernel<<<capped_grid(count, kThreads, 8), kThreads, 0, stream>>>(in, out, count, base);
    LAPPER_CUDA_CHECK(cudaGetLastError());
}

namespace {
// the per-warp emit kernel in its two shapes (one of them works, see FillShape)
void launch_fill_shapes(const FindView &f, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq, const uint64_t *d_offsets,
                        uint32_t *d_indices, uint32_t *heavy, unsigned int *heavy_count, const uint32_t *tile_list,
                        const unsigned int *tile_count, unsigned int *admitted, uint32_t max_admitted, cudaStream_t stream) {
    constexpr size_t kSparseBytes = (size_t)(kThreads / 32) * FillSparse::kWords * sizeof(uint32_t);
    constexpr size_t kDenseBytes = (size_t)(kThreads / 32) * FillDense::kWords * sizeof(uint32_t);
    static std::atomic<bool> prepared[64];
    int dev = 0;
    LAPPER_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !prepared[dev].load(std::memory_order_acquire)) {
        LAPPER_CUDA_CHECK(cudaFuncSetAttribute(find_fill_kernel<FillSparse>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSparseBytes));
        LAPPER_CUDA_CHECK(cudaFuncSetAttribute(find_fill_kernel<FillDense>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDenseBytes));
        prepared[dev].store(true, std::memory_order_release);
    }
    const uint32_t grid_s = capped_grid(nq, kThreads, FillSparse::kMinCtas * 8);
    const uint32_t grid_d = capped_grid(nq, kThreads, FillDense::kMinCtas * 8);
    g_query_launches++, find_fill_kernel<FillSparse><<<grid_s, kThreads, kSparseBytes, stream>>>(
        f, d_qs, d_qe, nq, d_offsets, d_indices, heavy, heavy_count, tile_list, tile_count, admitted, max_admitted);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    g_query_launches++, find_fill_kernel<FillDense><<<grid_d, kThreads, kDenseBytes, stream>>>(
        f, d_qs, d_qe, nq, d_offsets, d_indices, heavy, heavy_count, tile_list, tile_count, admitted, max_admitted);
    LAPPER_CUDA_CHECK(cudaGetLastError());
}

// the emit kernels over offsets that exist (tc == nullptr) or that the tile kernel derives from the counts (fused form)
void launch_emit(const FindView &f, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq, uint64_t *d_offsets,
                 uint32_t *d_indices, uint64_t heavy_bound, const TileCounts *tc, bool tiled, cudaStream_t stream) {
    // the heavy list: queries with >= kHeavyHits hits (at most heavy_bound / kHeavyHits) + light queries with long class
    // windows, admitted up to a sixteenth of the batch
    const uint32_t max_admitted = (uint32_t)std::min<uint64_t>(nq, nq / 16 + 1024);
    const uint64_t cap = std::min<uint64_t>(heavy_bound / kHeavyHits + 1, nq) + max_admitted;
    const uint64_t ntiles = div_up(nq, kFindTile);
    Scratch<uint32_t> heavy(cap, stream);
    Scratch<uint32_t> fallback(ntiles, stream);
    Scratch<unsigned int> counters(4, stream);  // heavy queries, next heavy query, fallback tiles, admitted long-window queries
    LAPPER_CUDA_CHECK(cudaMemsetAsync(counters.p, 0, 4 * sizeof(unsigned int), stream));
    if (tiled) {
        // tile form first; the tiles it cannot take (no locality) are left to the per-warp kernel
        // once per device: all of the SM's L1/shared array as shared memory, so that LAPPER_TILE_CTAS tiles are resident
        static std::atomic<bool> prepared[64];
        int dev = 0;
        LAPPER_CUDA_CHECK(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64 && !prepared[dev].load(std::memory_order_acquire)) {
            LAPPER_CUDA_CHECK(cudaFuncSetAttribute(find_fill_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   kTSmemWords * (int)sizeof(uint32_t)));
            LAPPER_CUDA_CHECK(cudaFuncSetAttribute(find_fill_tile_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                   cudaSharedmemCarveoutMaxShared));
            LAPPER_CUDA_CHECK(cudaFuncSetAttribute(find_fill_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   kTSmemWords * (int)sizeof(uint32_t)));
            LAPPER_CUDA_CHECK(cudaFuncSetAttribute(find_fill_tile_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                   cudaSharedmemCarveoutMaxShared));
            prepared[dev].store(true, std::memory_order_release);
        }
        const uint32_t tgrid = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)sm_count() * LAPPER_TILE_CTAS);
        if (tc)
            g_query_launches++, find_fill_tile_kernel<true><<<tgrid, kTT, kTSmemWords * sizeof(uint32_t), stream>>>(
                f, d_qs, d_qe, nq, d_offsets, *tc, d_indices, heavy.p, counters.p, fallback.p, counters.p + 2, counters.p + 3, max_admitted);
        else
            g_query_launches++, find_fill_tile_kernel<false><<<tgrid, kTT, kTSmemWords * sizeof(uint32_t), stream>>>(
                f, d_qs, d_qe, nq, d_offsets, TileCounts{}, d_indices, heavy.p, counters.p, fallback.p, counters.p + 2, counters.p + 3, max_admitted);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        launch_fill_shapes(f, d_qs, d_qe, nq, d_offsets, d_indices, heavy.p, counters.p, fallback.p, counters.p + 2, counters.p + 3,
                           max_admitted, stream);
    } else {
        launch_fill_shapes(f, d_qs, d_qe, nq, d_offsets, d_indices, heavy.p, counters.p, nullptr, nullptr, counters.p + 3, max_admitted,
                           stream);
    }
    // persistent warps; exits at once when the list is empty (it may also hold light queries with long windows)
    g_query_launches++, find_heavy_kernel<<<sm_count() * 8, kThreads, 0, stream>>>(f, d_qs, d_qe, d_offsets, d_indices, heavy.p, counters.p, counters.p + 1);
    LAPPER_CUDA_CHECK(cudaGetLastError());
}
}  // namespace

void launch_find_fill(const DeviceIndex &ix, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq,
                      const uint64_t *d_offsets, uint32_t *d_indices, cudaStream_t stream, uint64_t total) {
    if (nq == 0 || ix.n == 0) return;
    LAPPER_REQUIRE(nq < (1ull << 32), "find_batch: at most 2^32 - 1 queries per call");
    const FindView f = make_find_view(ix);
    // heavy queries (>= kHeavyHits hits): at most total / kHeavyHits of them; the total is offsets[nq]
    if (total == kTotalUnknown) {
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(&total, d_offsets + nq, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    const bool tiled = ((reinterpret_cast<uintptr_t>(d_offsets) | reinterpret_cast<uintptr_t>(d_qs) |
                         reinterpret_cast<uintptr_t>(d_qe)) & 15u) == 0 && !find_tile_disabled();
    launch_emit(f, d_qs, d_qe, nq, const_cast<uint64_t *>(d_offsets), d_indices, total, nullptr, tiled, stream);
}

// find_batch in one enqueue: count pass, scan of the tile sums, emit -- the tile kernel derives the offsets from the
// counts (no write_offsets pass, no host round trip between the phases).  Returns the total (synchronises the stream
// once, at the end); nothing is written to d_indices when it exceeds `capacity` (the offsets are complete either way).
uint64_t launch_find(const DeviceIndex &ix, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq, uint64_t *d_offsets,
                     uint32_t *d_indices, uint64_t capacity, cudaStream_t stream) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_offsets) | reinterpret_cast<uintptr_t>(d_qs) |
                           reinterpret_cast<uintptr_t>(d_qe)) & 31u) == 0;
    if (nq == 0 || ix.n == 0 || ix.has_inverted || !ix.buckets || nq >= (1ull << 31) || !aligned || find_tile_disabled()) {
        // shapes the fused form does not cover: the two phases, one after the other
        const uint64_t total = launch_find_offsets(ix, d_qs, d_qe, nq, d_offsets, stream);
        if (total <= capacity && total) launch_find_fill(ix, d_qs, d_qe, nq, d_offsets, d_indices, stream, total);
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
        return total;
    }
    const FindView f = make_find_view(ix);
    const uint64_t ntiles = div_up(nq, kFindTile), nsuper = div_up(ntiles, kSuperTiles);
    Scratch<uint32_t> counts(nq, stream);
    Scratch<uint64_t> tile_sums(ntiles, stream), tile_base(ntiles, stream), super_sums(nsuper, stream);
    const BucketPlan plan = plan_buckets(f.bv, d_qs, d_qe, nq, counts.p, true);
    LAPPER_CUDA_CHECK(cudaMemsetAsync(tile_sums.p, 0, ntiles * sizeof(uint64_t), stream));
    launch_count_buckets(ix, plan, d_qs, d_qe, nq, counts.p, nullptr, tile_sums.p, stream);
    g_query_launches++, super_sums_kernel<<<(uint32_t)nsuper, kSuperTiles, 0, stream>>>(tile_sums.p, ntiles, tile_base.p, super_sums.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    g_query_launches++, scan_super_sums_kernel<<<1, 1024, 0, stream>>>(super_sums.p, nsuper, d_offsets + nq);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    TileCounts tc;
    tc.counts = counts.p, tc.tile_sums = tile_sums.p, tc.tile_base = tile_base.p, tc.super_base = super_sums.p;
    tc.offsets_out = d_offsets, tc.capacity = capacity;
    launch_emit(f, d_qs, d_qe, nq, d_offsets, d_indices, capacity, &tc, true, stream);
    uint64_t total = 0;
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(&total, d_offsets + nq, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    return total;
}

}  // namespace lap
