// engine64.cu -- Lapper<I, T> with I = u64 (src/lib.rs:88-100: any PrimInt + Unsigned coordinate; the README's examples
// use usize).  The second instantiation of the query path: same C-ABI shape as the u32 engine (lapper64_* in
// include/lapper_b200.h), checked by the same CPU restatement built with I = u64 and the same parity suite
// (tests/test_gpu_u64.py).
//
// The u32 engine's tables (BITS sectors with 16-bit lanes, the packed 12-bit table, length classes) are formats for
// 32-bit coordinates in a GRCh38-sized universe.  64-bit coordinates get the plain form of the same algorithms:
//   build   (src/lib.rs:196-236)  two stable LSD radix sorts (by stop, then by start: (start, stop) order, ties in
//                                 input order like Vec::sort), sorted stops, max_len / inverted flag, prefix max of stops
//   count   (src/lib.rs:610-618)  lower_bound(starts, stop) - upper_bound(stops, start) by bisection under a
//                                 1024-entry sampled directory of each array in shared memory (the top ten levels)
//   find    (src/lib.rs:628-639, :695-718)  count pass (the BITS identity where it holds, the predicate elsewhere),
//                                 u64 exclusive scan, emit: one warp per query scans the window
//                                 [max(lower_bound(start - max_len), first i with prefix_max_i > start), lower_bound(stop))
//                                 32 intervals per step and ballot-compacts the hits -- ascending positions, the
//                                 reference iterator's order
//   cov, merge_overlaps, union_and_intersect (src/lib.rs:310-350, :361-416, :512-559)  the prefix-max closed forms of
//                                 setops.cu on u64 values
//   insert, depth, the seek cursor (src/lib.rs:260-273, :576-598 + :720-784, :658-671)  as in the u32 engine: sorted insert at
//                                 the reference's positions, depth in closed form over the sorted unique event coordinates
// At most 2^32 - 2 intervals (positions are u32, as in the u32 engine).  Replica groups exist for u32 only.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <vector>

#include "abi_common.cuh"
#include "index.cuh"
#include "merge_sequential.cuh"

struct lapper64 {
    int device = 0;
    uint64_t n = 0;
    uint64_t *S = nullptr;    // starts, sorted by (start, stop)
    uint64_t *Eiv = nullptr;  // stops paired with S
    uint64_t *E = nullptr;    // stops, sorted on their own (the reference's private `stops`)
    uint64_t *P = nullptr;    // prefix max of Eiv
    uint32_t *perm = nullptr; // sorted position -> input position
    uint64_t max_len = 0;
    bool has_inverted = false;
    bool overlaps_merged = false;
    bool cov_is_some = false;
    uint64_t cov = 0;
    uint64_t next_input_id = 0;  // input position handed to the next inserted interval
    uint64_t max_coord = 0;      // largest start or stop
    uint64_t min_coord = 0;      // smallest start or stop
    uint32_t query_len_hint = 0; // lapper64_tune_query_length: handed to the u32 engine of a narrowed set
    // The u32 engine over the same sorted intervals (see "narrowing" below): built on the first batch that is worth it
    // when every coordinate fits, dropped by insert / merge_overlaps.
    mutable std::atomic<lapper_t *> narrow{nullptr};
    mutable uint64_t narrow_base = 0, narrow_add = 0, narrow_top = 0xFFFFFFFEull;  // the coordinate map of `narrow` (NarrowMap)
    mutable std::atomic<bool> narrow_failed{false};  // building it failed once (memory): the 64-bit kernels answer from then on
    mutable std::mutex narrow_mu;
};

namespace lap {
namespace {

constexpr int kT = 256;
constexpr int kItems = 8;
constexpr int kTile64 = kT * kItems;  // items per tile of the scans

uint32_t grid_for(uint64_t n, int per_block = kT) { return (uint32_t)div_up(n ? n : 1, per_block); }
uint32_t capped(uint64_t n, int waves = 8) {
    const uint64_t g = div_up(n ? n : 1, kT), cap = (uint64_t)sm_count() * waves;
    return (uint32_t)std::min(g, cap);
}

__device__ __forceinline__ uint64_t lb64(const uint64_t *__restrict__ a, uint64_t lo, uint64_t hi, uint64_t key) {
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}
__device__ __forceinline__ uint64_t ub64(const uint64_t *__restrict__ a, uint64_t lo, uint64_t hi, uint64_t key) {
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) <= key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------ build
__global__ void iota_kernel(uint32_t *__restrict__ v, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}
__global__ void gather64_kernel(uint64_t *__restrict__ out, const uint64_t *__restrict__ in, const uint32_t *__restrict__ idx,
                                uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[idx[i]];
}
// max_len with checked_sub -> 0 (src/lib.rs:218-227), the inverted flag and the largest coordinate;
// out[0] = max_len, out[1] = inverted, out[2] = max(start, stop), out[3] = ~min(start, stop)
__global__ void __launch_bounds__(kT) stats64_kernel(const uint64_t *__restrict__ s, const uint64_t *__restrict__ e, uint64_t n,
                                                     unsigned long long *__restrict__ out) {
    unsigned long long mx = 0, inv = 0, top = 0, nlow = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t a = s[i], b = e[i];
        mx = max(mx, (unsigned long long)(b > a ? b - a : 0ull));
        inv |= (b < a) ? 1ull : 0ull;
        top = max(top, (unsigned long long)max(a, b));
        nlow = max(nlow, (unsigned long long)~min(a, b));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        inv |= __shfl_xor_sync(0xFFFFFFFFu, inv, o);
        top = max(top, __shfl_xor_sync(0xFFFFFFFFu, top, o));
        nlow = max(nlow, __shfl_xor_sync(0xFFFFFFFFu, nlow, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (mx) atomicMax(out, mx);
        if (inv) atomicOr(out + 1, 1ull);
        if (top) atomicMax(out + 2, top);
        if (nlow) atomicMax(out + 3, nlow);
    }
}

// ------------------------------------------------------------------------------------------ scans over u64 values
// Three kernels: per-tile aggregate, one block scans the aggregates (exclusive, in place; total out), per-tile scan
// with the carry applied, results handed to a sink functor as (index, inclusive, exclusive, value).
struct Max64 {
    static __device__ __forceinline__ uint64_t id() { return 0; }
    static __device__ __forceinline__ uint64_t op(uint64_t a, uint64_t b) { return a > b ? a : b; }
};
struct Add64 {
    static __device__ __forceinline__ uint64_t id() { return 0; }
    static __device__ __forceinline__ uint64_t op(uint64_t a, uint64_t b) { return a + b; }
};

template <typename Op>
__device__ __forceinline__ uint64_t block_scan_incl(uint64_t x, uint64_t *warp_tot, uint64_t &block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc = Op::op(y, inc);
    }
    __syncthreads();  // warp_tot may still be read from the previous use
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    uint64_t before = Op::id(), all = Op::id();
#pragma unroll
    for (int i = 0; i < kT / 32; i++) {
        if (i < w) before = Op::op(before, warp_tot[i]);
        all = Op::op(all, warp_tot[i]);
    }
    block_total = all;
    return Op::op(before, inc);
}

template <typename Op, typename In>
__global__ void __launch_bounds__(kT) tile_reduce64_kernel(In in, uint64_t n, uint64_t *__restrict__ agg) {
    __shared__ uint64_t warp_tot[kT / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kTile64 + (uint64_t)threadIdx.x * kItems;
    uint64_t x = Op::id();
#pragma unroll
    for (int k = 0; k < kItems; k++)
        if (base + k < n) x = Op::op(x, in(base + k));
    uint64_t total;
    block_scan_incl<Op>(x, warp_tot, total);
    if (threadIdx.x == 0) agg[blockIdx.x] = total;
}

template <typename Op>
__global__ void __launch_bounds__(kT) tile_scan64_kernel(uint64_t *__restrict__ agg, uint64_t ntiles, uint64_t *__restrict__ total_out) {
    __shared__ uint64_t warp_tot[kT / 32];
    uint64_t carry = Op::id();
    for (uint64_t t0 = 0; t0 < ntiles; t0 += kT) {
        const uint64_t t = t0 + threadIdx.x;
        const uint64_t x = t < ntiles ? agg[t] : Op::id();
        uint64_t total;
        const uint64_t inc = block_scan_incl<Op>(x, warp_tot, total);
        // exclusive = carry op (everything before t in this round)
        const uint64_t prev = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
        uint64_t ex;
        if ((threadIdx.x & 31) == 0) {
            ex = Op::id();
            for (int i = 0; i < (int)(threadIdx.x >> 5); i++) ex = Op::op(ex, warp_tot[i]);
        } else {
            ex = prev;
        }
        if (t < ntiles) agg[t] = Op::op(carry, ex);
        carry = Op::op(carry, total);
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename Op, typename In, typename Sink>
__global__ void __launch_bounds__(kT) tile_apply64_kernel(In in, uint64_t n, const uint64_t *__restrict__ tile_excl, Sink sink) {
    __shared__ uint64_t warp_tot[kT / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kTile64 + (uint64_t)threadIdx.x * kItems;
    uint64_t v[kItems], x = Op::id();
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        v[k] = base + k < n ? in(base + k) : Op::id();
        x = Op::op(x, v[k]);
    }
    uint64_t total;
    const uint64_t inc = block_scan_incl<Op>(x, warp_tot, total);
    // exclusive prefix of this thread's first item: everything before it in the tile, and the tiles before
    const uint64_t prev = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    uint64_t ex;
    if ((threadIdx.x & 31) == 0) {
        ex = Op::id();
        for (int i = 0; i < (int)(threadIdx.x >> 5); i++) ex = Op::op(ex, warp_tot[i]);
    } else {
        ex = prev;
    }
    uint64_t run = Op::op(tile_excl[blockIdx.x], ex);
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        const uint64_t incl = Op::op(run, v[k]);
        if (base + k < n) sink(base + k, incl, run, v[k]);
        run = incl;
    }
}

template <typename Op, typename In, typename Sink>
void scan64(In in, uint64_t n, Sink sink, uint64_t *d_total, cudaStream_t stream) {
    if (n == 0) {
        if (d_total) LAPPER_CUDA_CHECK(cudaMemsetAsync(d_total, 0, 8, stream));
        return;
    }
    const uint64_t ntiles = div_up(n, kTile64);
    Scratch<uint64_t> agg(ntiles, stream);
    tile_reduce64_kernel<Op, In><<<(uint32_t)ntiles, kT, 0, stream>>>(in, n, agg.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    tile_scan64_kernel<Op><<<1, kT, 0, stream>>>(agg.p, ntiles, d_total);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    tile_apply64_kernel<Op, In, Sink><<<(uint32_t)ntiles, kT, 0, stream>>>(in, n, agg.p, sink);
    LAPPER_CUDA_CHECK(cudaGetLastError());
}

struct Load64 {
    const uint64_t *a;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return a[i]; }
};
struct Load32 {
    const uint32_t *a;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return a[i]; }
};
struct StoreIncl {
    uint64_t *out;
    __device__ __forceinline__ void operator()(uint64_t i, uint64_t incl, uint64_t, uint64_t) const { out[i] = incl; }
};
struct StoreExcl {
    uint64_t *out;
    __device__ __forceinline__ void operator()(uint64_t i, uint64_t, uint64_t excl, uint64_t) const { out[i] = excl; }
};
// merge_overlaps (src/lib.rs:364-381): a run starts where the running max of the stops before it is below the start
struct Head64 {
    const uint64_t *S, *P;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return (i == 0 || P[i - 1] < S[i]) ? 1ull : 0ull; }
};
struct MergeSink64 {
    const uint64_t *S, *P;
    const uint32_t *perm;
    uint64_t n;
    uint64_t *ms, *me;
    uint32_t *mperm;
    __device__ __forceinline__ void operator()(uint64_t i, uint64_t incl, uint64_t, uint64_t flag) const {
        if (flag) ms[incl - 1] = S[i], mperm[incl - 1] = perm[i];  // the head gives start and val (src/lib.rs:383-390)
        if (i + 1 == n || P[i] < S[i + 1]) me[incl - 1] = P[i];    // the run's last interval: its stop is the running max
    }
};
struct RunLen64 {
    const uint64_t *s, *e;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return e[i] - s[i]; }
};

// ------------------------------------------------------------------------------------------ count
constexpr int kDir = 1024;  // sampled keys per array in shared memory: the top ten levels of each bisection

struct View64 {
    const uint64_t *S, *Eiv, *E, *P;
    uint64_t n, max_len;
    uint64_t stride;  // directory sample i is element min(i * stride, n - 1)
    bool has_inverted;
};

__device__ __forceinline__ void load_dir(const View64 &v, uint64_t *dS, uint64_t *dE) {
    for (int i = threadIdx.x; i < kDir; i += blockDim.x) {
        const uint64_t at = min((uint64_t)i * v.stride, v.n - 1);
        dS[i] = v.S[at];
        dE[i] = v.E[at];
    }
    __syncthreads();
}
// #{x in a : x < key} with the directory narrowing the range first
__device__ __forceinline__ uint64_t dir_lb(const uint64_t *__restrict__ a, const uint64_t *dir, const View64 &v, uint64_t key) {
    // samples below key: all elements up to the last such sample are below key
    int lo = 0, hi = kDir;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (dir[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    // lo samples are < key: elements [0, (lo - 1) * stride] are < key; the answer is at most lo * stride
    const uint64_t from = lo ? min((uint64_t)(lo - 1) * v.stride + 1, v.n) : 0;
    const uint64_t to = min((uint64_t)lo * v.stride, v.n);
    return lb64(a, from, to, key);
}
__device__ __forceinline__ uint64_t dir_ub(const uint64_t *__restrict__ a, const uint64_t *dir, const View64 &v, uint64_t key) {
    int lo = 0, hi = kDir;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (dir[mid] <= key)
            lo = mid + 1;
        else
            hi = mid;
    }
    const uint64_t from = lo ? min((uint64_t)(lo - 1) * v.stride + 1, v.n) : 0;
    const uint64_t to = min((uint64_t)lo * v.stride, v.n);
    return ub64(a, from, to, key);
}

// Lapper::count (src/lib.rs:611-617): len - first - (len - last) in wrapping usize == last - first (mod 2^64);
// `start + 1` wraps to 0 at I::MAX in a release build, so first = 0 there.
__global__ void __launch_bounds__(kT) count64_kernel(const View64 v, const uint64_t *__restrict__ qs, const uint64_t *__restrict__ qe,
                                                     uint64_t nq, uint64_t *__restrict__ out) {
    __shared__ uint64_t dS[kDir], dE[kDir];
    load_dir(v, dS, dE);
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nq; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t a = qs[j], b = qe[j];
        const uint64_t last = dir_lb(v.S, dS, v, b);
        const uint64_t first = a == ~0ull ? 0ull : dir_ub(v.E, dE, v, a);
        out[j] = last - first;
    }
}

// ------------------------------------------------------------------------------------------ find
// the scan window of one query: the reference's [lower_bound(start - max_len), first interval with start >= stop)
// (src/lib.rs:632-635, :712-713), its front tightened by the running max of the stops (nothing before the first
// position whose prefix max exceeds `start` can end after `start`)
__device__ __forceinline__ void window64(const View64 &v, uint64_t a, uint64_t b, uint64_t &lo, uint64_t &hi) {
    hi = lb64(v.S, 0, v.n, b);
    const uint64_t floor_key = a >= v.max_len ? a - v.max_len : 0ull;
    lo = lb64(v.S, 0, hi, floor_key);
    const uint64_t tight = ub64(v.P, 0, hi, a);
    lo = max(lo, tight);
    if (lo > hi) lo = hi;
}

// number of positions find(a, b) yields
__global__ void __launch_bounds__(kT) find64_count_kernel(const View64 v, const uint64_t *__restrict__ qs,
                                                          const uint64_t *__restrict__ qe, uint64_t nq, uint32_t *__restrict__ counts) {
    __shared__ uint64_t dS[kDir], dE[kDir];
    load_dir(v, dS, dE);
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nq; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t a = qs[j], b = qe[j];
        uint32_t c;
        if (a < b && !v.has_inverted) {
            // |find| == count when every interval has start <= stop and the query is non-empty
            c = (uint32_t)(dir_lb(v.S, dS, v, b) - dir_ub(v.E, dE, v, a));
        } else {
            uint64_t lo, hi;
            window64(v, a, b, lo, hi);
            c = 0;
            for (uint64_t i = lo; i < hi; i++) c += __ldg(v.Eiv + i) > a;
        }
        counts[j] = c;
    }
}

// one warp per query: 32 intervals of the window per step, hits ballot-compacted into the query's slice of `indices`
__global__ void __launch_bounds__(kT) find64_emit_kernel(const View64 v, const uint64_t *__restrict__ qs, const uint64_t *__restrict__ qe,
                                                         uint64_t nq, const uint64_t *__restrict__ offsets, uint32_t *__restrict__ indices,
                                                         uint64_t capacity) {
    if (offsets[nq] > capacity) return;  // nothing is written when the result does not fit
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t nwarps = (uint64_t)gridDim.x * (kT / 32);
    for (uint64_t j = (uint64_t)blockIdx.x * (kT / 32) + (threadIdx.x >> 5); j < nq; j += nwarps) {
        const uint64_t o0 = offsets[j], room = offsets[j + 1] - o0;
        if (room == 0) continue;
        const uint64_t a = __ldg(qs + j), b = __ldg(qe + j);
        uint64_t lo, hi;
        window64(v, a, b, lo, hi);
        uint64_t written = 0;
        for (uint64_t i0 = lo; i0 < hi && written < room; i0 += 32) {
            const uint64_t i = i0 + lane;
            const bool hit = i < hi && __ldg(v.Eiv + i) > a;  // start_i < b holds for every i < hi (src/lib.rs:140-142)
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
            const uint64_t at = written + __popc(m & ((1u << lane) - 1u));
            if (hit && at < room) indices[o0 + at] = (uint32_t)i;
            written += __popc(m);
        }
    }
}

// ------------------------------------------------------------------------------------------ set operations
// cov = sum_{last of run} P_i - sum_{head} S_i (wrapping u64), runs of calculate_coverage (src/lib.rs:327-350)
__global__ void __launch_bounds__(kT) cov64_kernel(const uint64_t *__restrict__ S, const uint64_t *__restrict__ P, uint64_t n,
                                                   unsigned long long *__restrict__ out) {
    unsigned long long acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = S[i], p = P[i];
        if (i == 0 || P[i - 1] < s) acc -= s;
        if (i + 1 == n || p < S[i + 1]) acc += p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// covered length of union(B) below x; B given as disjoint sorted runs with exclusive prefix lengths
__device__ __forceinline__ uint64_t covered_below64(const uint64_t *__restrict__ bs, const uint64_t *__restrict__ be,
                                                    const uint64_t *__restrict__ cum, uint64_t m, uint64_t x) {
    const uint64_t j = lb64(bs, 0, m, x);  // runs starting below x
    if (j == 0) return 0;
    const uint64_t s = bs[j - 1], e = be[j - 1];
    return cum[j - 1] + (min(x, e) - s);
}
__global__ void __launch_bounds__(kT) intersect64_kernel(const uint64_t *__restrict__ as, const uint64_t *__restrict__ ae, uint64_t ma,
                                                         const uint64_t *__restrict__ bs, const uint64_t *__restrict__ be,
                                                         const uint64_t *__restrict__ cum, uint64_t mb,
                                                         unsigned long long *__restrict__ out) {
    unsigned long long acc = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < ma; r += (uint64_t)gridDim.x * blockDim.x)
        acc += covered_below64(bs, be, cum, mb, ae[r]) - covered_below64(bs, be, cum, mb, as[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// ------------------------------------------------------------------------------------------ depth, insert, seek
// depth() in closed form (the same one as setops.cu, see compute_depth there): over the sorted unique event
// coordinates u_t with D_t = #{start <= u_t} - #{stop <= u_t}, a run starts wherever D changes to a positive value and
// ends at the next change; a zero-length merged run is reported as (u, u, 0) unless it is the first run of the set.
struct Unique64 {
    const uint64_t *X;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return (i == 0 || X[i] != X[i - 1]) ? 1ull : 0ull; }
};
struct UniqueSink64 {
    const uint64_t *X;
    uint64_t *U;
    __device__ __forceinline__ void operator()(uint64_t i, uint64_t incl, uint64_t, uint64_t flag) const {
        if (flag) U[incl - 1] = X[i];
    }
};
__global__ void __launch_bounds__(kT) depth_eval64_kernel(const uint64_t *__restrict__ S, const uint64_t *__restrict__ E, uint64_t n,
                                                          const uint64_t *__restrict__ U, uint64_t m, uint32_t *__restrict__ D,
                                                          uint32_t *__restrict__ has_start) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t u = U[t];
        const uint64_t ub_s = ub64(S, 0, n, u);
        D[t] = (uint32_t)(ub_s - ub64(E, 0, n, u));
        has_start[t] = lb64(S, 0, n, u) < ub_s ? 1u : 0u;
    }
}
struct EmitPoint64 {
    const uint32_t *D, *has_start;
    __device__ __forceinline__ bool zero_run(uint64_t t) const { return t > 0 && D[t] == 0 && D[t - 1] == 0 && has_start[t]; }
    __device__ __forceinline__ uint64_t operator()(uint64_t t) const {
        return (t == 0 || D[t] != D[t - 1] || zero_run(t)) ? 1ull : 0ull;
    }
};
struct EmitPointSink64 {
    const uint64_t *U;
    EmitPoint64 f;
    uint64_t *eu;
    uint32_t *ed, *ez;
    __device__ __forceinline__ void operator()(uint64_t t, uint64_t incl, uint64_t, uint64_t flag) const {
        if (flag) eu[incl - 1] = U[t], ed[incl - 1] = f.D[t], ez[incl - 1] = f.zero_run(t) ? 1u : 0u;
    }
};
struct Run64 {
    const uint32_t *ed, *ez;
    __device__ __forceinline__ uint64_t operator()(uint64_t q) const { return (ed[q] > 0 || ez[q]) ? 1ull : 0ull; }
};
struct RunSink64 {
    const uint64_t *eu;
    const uint32_t *ed, *ez;
    uint64_t *os, *oe;
    uint32_t *od;
    __device__ __forceinline__ void operator()(uint64_t q, uint64_t incl, uint64_t, uint64_t flag) const {
        if (flag) {
            const uint64_t r = incl - 1;
            os[r] = eu[q];
            oe[r] = ez[q] ? eu[q] : eu[q + 1];  // a positive-depth run ends at the next emit point, always a depth change
            od[r] = ed[q];
        }
    }
};

// positions the reference's insert() picks (src/lib.rs:261-263): lower_bound of the interval among the sorted
// intervals by (start, stop), and of its stop among the sorted stops
__global__ void insert_positions64_kernel(const uint64_t *__restrict__ S, const uint64_t *__restrict__ Eiv,
                                          const uint64_t *__restrict__ E, uint64_t n, uint64_t start, uint64_t stop,
                                          uint64_t *__restrict__ out) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        const uint64_t s = S[mid], e = Eiv[mid];
        if (s < start || (s == start && e < stop))
            lo = mid + 1;
        else
            hi = mid;
    }
    out[0] = lo;
    out[1] = lb64(E, 0, n, stop);
}
template <typename T>
__global__ void insert_into64_kernel(T *__restrict__ dst, const T *__restrict__ src, uint64_t n, uint64_t pos, T value) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = i < pos ? src[i] : i == pos ? value : src[i - 1];
}
// Lapper::seek's cursor update (src/lib.rs:658-671) with the advancing while-loop in closed form
__global__ void seek_cursor64_kernel(const uint64_t *__restrict__ S, uint64_t n, uint64_t max_len, uint64_t start,
                                     uint64_t cursor_in, uint64_t *__restrict__ cursor_out) {
    const uint64_t floor_key = start >= max_len ? start - max_len : 0ull;
    uint64_t c = cursor_in;
    if (c == 0 || (c < n && S[c] > start)) c = lb64(S, 0, n, floor_key);
    if (c + 1 < n && S[c + 1] < floor_key) c = lb64(S, 0, n, floor_key) - 1;
    *cursor_out = c;
}

// ------------------------------------------------------------------------------------------ host side
View64 view_of(const lapper64 *l) {
    View64 v;
    v.S = l->S, v.Eiv = l->Eiv, v.E = l->E, v.P = l->P;
    v.n = l->n, v.max_len = l->max_len;
    v.stride = std::max<uint64_t>(div_up(l->n ? l->n : 1, kDir), 1);
    v.has_inverted = l->has_inverted;
    return v;
}

void free_arrays(lapper64 *l, cudaStream_t stream) {
    pool_free(l->S, stream), pool_free(l->Eiv, stream), pool_free(l->E, stream), pool_free(l->P, stream), pool_free(l->perm, stream);
    l->S = l->Eiv = l->E = l->P = nullptr;
    l->perm = nullptr;
}

// max_len, inverted flag and prefix max from the sorted arrays (S, Eiv set; P allocated here)
void finish_from_sorted(lapper64 *l, cudaStream_t stream) {
    uint64_t bytes = 0;
    l->max_len = 0, l->has_inverted = false, l->max_coord = 0, l->min_coord = 0;
    if (l->n == 0) return;
    l->P = pool_alloc<uint64_t>(l->n, bytes, stream);
    Scratch<unsigned long long> st(4, stream);
    LAPPER_CUDA_CHECK(cudaMemsetAsync(st.p, 0, 32, stream));
    stats64_kernel<<<capped(l->n), kT, 0, stream>>>(l->S, l->Eiv, l->n, st.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    scan64<Max64>(Load64{l->Eiv}, l->n, StoreIncl{l->P}, nullptr, stream);
    unsigned long long h[4] = {0, 0, 0, 0};
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(h, st.p, 32, cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    l->max_len = h[0];
    l->has_inverted = h[1] != 0;
    l->max_coord = h[2];
    l->min_coord = ~h[3];
}

void build_from_device(lapper64 *l, const uint64_t *d_s, const uint64_t *d_e, uint64_t n, cudaStream_t stream) {
    LAPPER_REQUIRE(n < 0xFFFFFFFFull, "lapper64: at most 2^32 - 2 intervals (positions are u32)");
    l->n = n;
    if (n == 0) return;
    uint64_t bytes = 0;
    l->S = pool_alloc<uint64_t>(n, bytes, stream);
    l->Eiv = pool_alloc<uint64_t>(n, bytes, stream);
    l->E = pool_alloc<uint64_t>(n, bytes, stream);
    l->perm = pool_alloc<uint32_t>(n, bytes, stream);
    Scratch<uint64_t> ktmp(n, stream);
    Scratch<uint32_t> vtmp(n, stream);
    // stable sort by stop, then stable sort by start: (start, stop) order, equal intervals in input order
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(l->E, d_e, n * 8, cudaMemcpyDeviceToDevice, stream));
    iota_kernel<<<grid_for(n), kT, 0, stream>>>(l->perm, n);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    launch_sort_pairs_u64(l->E, ktmp.p, l->perm, vtmp.p, n, stream);  // E: stops ascending (= the sorted stops), perm: their inputs
    gather64_kernel<<<grid_for(n), kT, 0, stream>>>(l->S, d_s, l->perm, n);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    launch_sort_pairs_u64(l->S, ktmp.p, l->perm, vtmp.p, n, stream);
    gather64_kernel<<<grid_for(n), kT, 0, stream>>>(l->Eiv, d_e, l->perm, n);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    finish_from_sorted(l, stream);
}

uint64_t cov_of(const lapper64 *l, cudaStream_t stream) {
    if (l->n == 0) return 0;
    Scratch<unsigned long long> d(1, stream);
    LAPPER_CUDA_CHECK(cudaMemsetAsync(d.p, 0, 8, stream));
    cov64_kernel<<<capped(l->n), kT, 0, stream>>>(l->S, l->P, l->n, d.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    unsigned long long h = 0;
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(&h, d.p, 8, cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    return h;
}

struct Merged64 {
    uint64_t m = 0;
    uint64_t *s = nullptr, *e = nullptr;
    uint32_t *perm = nullptr;
    cudaStream_t stream;
    explicit Merged64(cudaStream_t st) : stream(st) {}
    ~Merged64() { pool_free(s, stream), pool_free(e, stream), pool_free(perm, stream); }
    Merged64(const Merged64 &) = delete;
    Merged64 &operator=(const Merged64 &) = delete;
};

void merged_of(const lapper64 *l, Merged64 &out, cudaStream_t stream) {
    if (l->n == 0) return;
    if (l->has_inverted) {
        // some stop < start: the reference's own loop, one warp (merge_sequential.cuh)
        Scratch<unsigned long long> d_m(1, stream);
        merge_sequential_kernel<uint64_t, false><<<1, 32, 0, stream>>>(l->S, l->Eiv, l->perm, l->n, nullptr, nullptr, nullptr, d_m.p);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        unsigned long long m = 0;
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(&m, d_m.p, 8, cudaMemcpyDeviceToHost, stream));
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
        uint64_t bytes = 0;
        out.m = m;
        out.s = pool_alloc<uint64_t>(m, bytes, stream);
        out.e = pool_alloc<uint64_t>(m, bytes, stream);
        out.perm = pool_alloc<uint32_t>(m, bytes, stream);
        merge_sequential_kernel<uint64_t, true><<<1, 32, 0, stream>>>(l->S, l->Eiv, l->perm, l->n, out.s, out.e, out.perm, nullptr);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        return;
    }
    const uint64_t n = l->n, ntiles = div_up(n, kTile64);
    const Head64 flags{l->S, l->P};
    Scratch<uint64_t> agg(ntiles, stream), d_total(1, stream);
    tile_reduce64_kernel<Add64, Head64><<<(uint32_t)ntiles, kT, 0, stream>>>(flags, n, agg.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    tile_scan64_kernel<Add64><<<1, kT, 0, stream>>>(agg.p, ntiles, d_total.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    uint64_t m = 0;
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(&m, d_total.p, 8, cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    uint64_t bytes = 0;
    out.m = m;
    out.s = pool_alloc<uint64_t>(m, bytes, stream);
    out.e = pool_alloc<uint64_t>(m, bytes, stream);
    out.perm = pool_alloc<uint32_t>(m, bytes, stream);
    tile_apply64_kernel<Add64, Head64, MergeSink64><<<(uint32_t)ntiles, kT, 0, stream>>>(
        flags, n, agg.p, MergeSink64{l->S, l->P, l->perm, n, out.s, out.e, out.perm});
    LAPPER_CUDA_CHECK(cudaGetLastError());
}

void need_valid(const lapper64 *l, const char *what) {
    if (l->has_inverted)
        throw Error(LAPPER_ERR_PRECONDITION, std::string(what) + ": an interval has stop < start (the reference's arithmetic is "
                                                                "undefined there; the closed form needs start <= stop)");
}

// offsets[0 .. nq] from the count pass; returns nothing (the total is offsets[nq])
void find_offsets64(const lapper64 *l, const uint64_t *d_qs, const uint64_t *d_qe, uint64_t nq, uint64_t *d_offsets, cudaStream_t stream) {
    if (nq == 0 || l->n == 0) {
        LAPPER_CUDA_CHECK(cudaMemsetAsync(d_offsets, 0, (nq + 1) * 8, stream));
        return;
    }
    Scratch<uint32_t> counts(nq, stream);
    find64_count_kernel<<<capped(nq), kT, 0, stream>>>(view_of(l), d_qs, d_qe, nq, counts.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    scan64<Add64>(Load32{counts.p}, nq, StoreExcl{d_offsets}, d_offsets + nq, stream);
}

void find_emit64(const lapper64 *l, const uint64_t *d_qs, const uint64_t *d_qe, uint64_t nq, const uint64_t *d_offsets,
                 uint32_t *d_indices, uint64_t capacity, cudaStream_t stream) {
    if (nq == 0 || l->n == 0) return;
    find64_emit_kernel<<<(uint32_t)std::min<uint64_t>(div_up(nq, kT / 32), (uint64_t)sm_count() * 16), kT, 0, stream>>>(
        view_of(l), d_qs, d_qe, nq, d_offsets, d_indices, capacity);
    LAPPER_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------ narrowing
// A Lapper<u64, T> -- the README's `usize` -- whose coordinates all fit 32 bits (every genome does) is served by the u32
// engine's tables and kernels: the sorted arrays are narrowed once into a u32 index (positions and tie order are those
// of this handle: lapper_from_sorted_dev keeps the order it is given), and a batch's queries are narrowed on the way in:
//     stop'  = min(stop, 2^32 - 2)
//     start' = start == u64::MAX ? u32::MAX : min(start, 2^32 - 2)
// With every coordinate <= 2^32 - 3 each comparison the reference makes (start_i < stop, stop_i <= start resp.
// stop_i > start; src/lib.rs:140-142, :611-618) has the same outcome on the narrowed query, and `start + 1` wraps for
// start' = u32::MAX exactly as it does for start = u64::MAX (:614).  Counts are u64 either way, positions u32.
// The u64 arrays stay: the set operations, insert, depth and the seek cursor work on them.
//
// Rebased form: a set with coordinates beyond 32 bits whose SPAN fits (max - min <= 2^32 - 4: nanosecond timestamps
// within four seconds, microseconds within 71 minutes, seconds at any epoch) is narrowed relative to its smallest
// coordinate.  With base = min, R = max - min:   x' = x - base + 1 for the set's coordinates (1 .. R + 1), and for a
// query coordinate q:   q' = 0 if q < base,   q - base + 1 if q - base <= R,   R + 2 beyond;   start' = u32::MAX for
// start = u64::MAX as above.  The map is monotone and sends "below every coordinate" / "above every coordinate" to
// values below / above every narrowed coordinate, so x < q, x <= q and x > q keep their outcomes for every x in the set.
std::atomic<int> g_narrow_mode{1};  // 0 = never, 1 = from the first batch of kNarrowMinBatch queries on, 2 = always
constexpr uint64_t kNarrowMaxCoord = 0xFFFFFFFDull;
constexpr uint64_t kNarrowMaxSpan = 0xFFFFFFFCull;
constexpr uint64_t kNarrowMinBatch = 4096;

// q' = q < base ? 0 : min(q - base + add, top); plain form: base = 0, add = 0, top = 2^32 - 2
struct NarrowMap {
    uint64_t base, add, top;
    __host__ __device__ uint32_t operator()(uint64_t q) const {
        if (q < base) return 0u;
        const uint64_t d = q - base;
        return (uint32_t)(d >= top - add ? top : d + add);
    }
};

// the map of a set, or false when neither form applies
bool narrow_map_of(const lapper64 *l, NarrowMap &m) {
    if (l->max_coord <= kNarrowMaxCoord) {
        m = NarrowMap{0, 0, 0xFFFFFFFEull};
        return true;
    }
    if (l->max_coord - l->min_coord <= kNarrowMaxSpan) {
        m = NarrowMap{l->min_coord, 1, l->max_coord - l->min_coord + 2};
        return true;
    }
    return false;
}

__global__ void __launch_bounds__(kT) narrow_index_kernel(const uint64_t *__restrict__ S, const uint64_t *__restrict__ Eiv,
                                                          const uint64_t *__restrict__ E, uint64_t n, uint32_t *__restrict__ s,
                                                          uint32_t *__restrict__ eiv, uint32_t *__restrict__ e, NarrowMap m) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        s[i] = m(S[i]), eiv[i] = m(Eiv[i]), e[i] = m(E[i]);
}

__global__ void __launch_bounds__(kT) narrow_queries_kernel(const uint64_t *__restrict__ qs, const uint64_t *__restrict__ qe,
                                                            uint64_t nq, uint32_t *__restrict__ a, uint32_t *__restrict__ b, NarrowMap m) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nq; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t x = qs[j], y = qe[j];
        a[j] = x == ~0ull ? 0xFFFFFFFFu : m(x);
        b[j] = m(y);
    }
}

// the u32 engine of this handle, or nullptr when the batch takes the 64-bit kernels
lapper_t *narrow_of(const lapper64 *l, uint64_t nq, cudaStream_t stream) {
    const int mode = g_narrow_mode.load(std::memory_order_relaxed);
    NarrowMap m;
    if (mode == 0 || l->n == 0 || l->n >= 0xFFFFFFF0ull || !narrow_map_of(l, m)) return nullptr;
    if (l->narrow_failed.load(std::memory_order_relaxed)) return nullptr;
    lapper_t *h = l->narrow.load(std::memory_order_acquire);
    if (h) return h;
    if (mode == 1 && nq < kNarrowMinBatch) return nullptr;
    std::lock_guard<std::mutex> guard(l->narrow_mu);
    h = l->narrow.load(std::memory_order_acquire);
    if (h) return h;
    // The u32 index is an accelerator, not a requirement: if it cannot be built (about 50 bytes per interval on top of
    // the u64 arrays), the batch and every later one go to the 64-bit kernels.
    try {
        Scratch<uint32_t> s(l->n, stream), eiv(l->n, stream), e(l->n, stream);
        narrow_index_kernel<<<capped(l->n), kT, 0, stream>>>(l->S, l->Eiv, l->E, l->n, s.p, eiv.p, e.p, m);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        const int rc = lapper_from_sorted_dev(s.p, eiv.p, e.p, l->perm, l->n, l->overlaps_merged ? 1 : 0, l->device, 0, stream, &h);
        if (rc != LAPPER_OK) throw Error(rc, last_error_slot());
        if (l->query_len_hint && lapper_tune_query_length(h, l->query_len_hint) != LAPPER_OK) {
            lapper_free(h);
            h = nullptr;
            throw Error(LAPPER_ERR_OOM, last_error_slot());
        }
    } catch (const Error &err) {
        if (err.code != LAPPER_ERR_OOM) throw;
        cudaGetLastError();
        l->narrow_failed.store(true, std::memory_order_relaxed);
        return nullptr;
    }
    l->narrow_base = m.base, l->narrow_add = m.add, l->narrow_top = m.top;  // (published by the release store below)
    l->narrow.store(h, std::memory_order_release);
    return h;
}

void drop_narrow(lapper64 *l) {
    l->narrow_failed.store(false, std::memory_order_relaxed);
    lapper_t *h = l->narrow.exchange(nullptr, std::memory_order_acq_rel);
    if (h) lapper_free(h);
}

struct NarrowQueries {
    Scratch<uint32_t> a, b;
    NarrowQueries(const lapper64 *l, const uint64_t *d_qs, const uint64_t *d_qe, uint64_t nq, cudaStream_t stream) : a(nq, stream), b(nq, stream) {
        if (nq == 0) return;
        const NarrowMap m{l->narrow_base, l->narrow_add, l->narrow_top};  // read after the acquire load of l->narrow
        narrow_queries_kernel<<<capped(nq), kT, 0, stream>>>(d_qs, d_qe, nq, a.p, b.p, m);
        LAPPER_CUDA_CHECK(cudaGetLastError());
    }
};

void check64(const lapper64 *l) { LAPPER_REQUIRE(l != nullptr, "null lapper64 handle"); }

}  // namespace
}  // namespace lap

using namespace lap;

extern "C" {

int lapper64_build_dev(const uint64_t *d_starts, const uint64_t *d_stops, uint64_t n, int device, void *stream, lapper64_t **out) {
    return guarded_call([&] {
        LAPPER_REQUIRE(out, "null output handle");
        *out = nullptr;
        LAPPER_REQUIRE(n == 0 || (d_starts && d_stops), "lapper64_build: null pointer");
        require_device_index(device);
        DeviceGuard dg(device);
        std::unique_ptr<lapper64> l(new lapper64());
        l->device = device;
        try {
            build_from_device(l.get(), d_starts, d_stops, n, (cudaStream_t)stream);
            l->next_input_id = n;
            LAPPER_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
        } catch (...) {
            free_arrays(l.get(), (cudaStream_t)stream);
            throw;
        }
        *out = l.release();
    });
}

int lapper64_build(const uint64_t *starts, const uint64_t *stops, uint64_t n, int device, lapper64_t **out) {
    return guarded_call([&] {
        LAPPER_REQUIRE(out, "null output handle");
        *out = nullptr;
        LAPPER_REQUIRE(n == 0 || (starts && stops), "lapper64_build: null pointer");
        require_device_index(device);
        DeviceGuard dg(device);
        OwnedStream st;
        Scratch<uint64_t> ds(n, st.s), de(n, st.s);
        if (n) {
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(ds.p, starts, n * 8, cudaMemcpyHostToDevice, st.s));
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(de.p, stops, n * 8, cudaMemcpyHostToDevice, st.s));
        }
        std::unique_ptr<lapper64> l(new lapper64());
        l->device = device;
        try {
            build_from_device(l.get(), ds.p, de.p, n, st.s);
            l->next_input_id = n;
            ds.release(), de.release();
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
        } catch (...) {
            free_arrays(l.get(), st.s);
            cudaStreamSynchronize(st.s);
            throw;
        }
        *out = l.release();
    });
}

void lapper64_free(lapper64_t *l) {
    if (!l) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(l->device);
    drop_narrow(l);
    free_arrays(l, nullptr);
    cudaStreamSynchronize(nullptr);
    if (prev >= 0) cudaSetDevice(prev);
    delete l;
}

uint64_t lapper64_len(const lapper64_t *l) { return l ? l->n : 0; }
uint64_t lapper64_max_len(const lapper64_t *l) { return l ? l->max_len : 0; }
int lapper64_overlaps_merged(const lapper64_t *l) { return l && l->overlaps_merged; }
int lapper64_has_inverted(const lapper64_t *l) { return l && l->has_inverted; }

int lapper64_get_intervals(const lapper64_t *l, uint64_t *starts_out, uint64_t *stops_out, uint32_t *perm_out) {
    return guarded_call([&] {
        check64(l);
        if (l->n == 0) return;
        DeviceGuard dg(l->device);
        if (starts_out) LAPPER_CUDA_CHECK(cudaMemcpy(starts_out, l->S, l->n * 8, cudaMemcpyDeviceToHost));
        if (stops_out) LAPPER_CUDA_CHECK(cudaMemcpy(stops_out, l->Eiv, l->n * 8, cudaMemcpyDeviceToHost));
        if (perm_out) LAPPER_CUDA_CHECK(cudaMemcpy(perm_out, l->perm, l->n * 4, cudaMemcpyDeviceToHost));
    });
}

int lapper64_get_sorted_stops(const lapper64_t *l, uint64_t *stops_sorted_out) {
    return guarded_call([&] {
        check64(l);
        LAPPER_REQUIRE(stops_sorted_out || l->n == 0, "null pointer");
        DeviceGuard dg(l->device);
        if (l->n) LAPPER_CUDA_CHECK(cudaMemcpy(stops_sorted_out, l->E, l->n * 8, cudaMemcpyDeviceToHost));
    });
}

int lapper64_count_batch_dev(const lapper64_t *l, const uint64_t *d_q_start, const uint64_t *d_q_stop, uint64_t nq,
                             uint64_t *d_counts_out, void *stream) {
    return guarded_call([&] {
        check64(l);
        LAPPER_REQUIRE(nq == 0 || (d_q_start && d_q_stop && d_counts_out), "count_batch: null pointer");
        if (nq == 0) return;
        DeviceGuard dg(l->device);
        if (l->n == 0) {
            LAPPER_CUDA_CHECK(cudaMemsetAsync(d_counts_out, 0, nq * 8, (cudaStream_t)stream));
            return;
        }
        if (lapper_t *h = narrow_of(l, nq, (cudaStream_t)stream)) {
            NarrowQueries q(l, d_q_start, d_q_stop, nq, (cudaStream_t)stream);
            const int rc = lapper_count_batch_u32_dev64(h, q.a.p, q.b.p, nq, d_counts_out, stream);
            if (rc != LAPPER_OK) throw Error(rc, last_error_slot());
            return;
        }
        g_query_launches++, count64_kernel<<<capped(nq), kT, 0, (cudaStream_t)stream>>>(view_of(l), d_q_start, d_q_stop, nq, d_counts_out);
        LAPPER_CUDA_CHECK(cudaGetLastError());
    });
}

int lapper64_count_batch(const lapper64_t *l, const uint64_t *q_start, const uint64_t *q_stop, uint64_t nq, uint64_t *counts_out) {
    return guarded_call([&] {
        check64(l);
        LAPPER_REQUIRE(nq == 0 || (q_start && q_stop && counts_out), "count_batch: null pointer");
        if (nq == 0) return;
        DeviceGuard dg(l->device);
        // chunks of 4 M queries on three streams: chunk k's counts go out while chunk k+1 is counted and chunk k+2's
        // queries come in (as the u32 engine's host path does; pinned buffers give the link's full rate both ways)
        constexpr int kSlots = 3;
        const uint64_t chunk = std::min<uint64_t>(nq, 1ull << 22);
        const int used = (int)std::min<uint64_t>(kSlots, div_up(nq, chunk));
        if (nq >= (4ull << 20) && host_copy_threads() > 0 &&
            !(host_pointer_is_pinned(q_start) && host_pointer_is_pinned(q_stop) && host_pointer_is_pinned(counts_out))) {
            // pageable buffers (a Vec<usize>): the same pipeline through the library's pinned staging buffers, filled and
            // drained by the copy threads (api.cu count_host_staged is the u32 twin)
            OwnedStream sst[kSlots];
            Scratch<uint64_t> dqs[kSlots], dqe[kSlots], dout[kSlots];
            std::vector<PinnedBuf> hqs, hqe, hout;
            uint64_t pending_lo[kSlots] = {}, pending_m[kSlots] = {};
            for (int s = 0; s < used; s++) {
                dqs[s].alloc(chunk, sst[s].s), dqe[s].alloc(chunk, sst[s].s), dout[s].alloc(chunk, sst[s].s);
                hqs.emplace_back(chunk * 8), hqe.emplace_back(chunk * 8), hout.emplace_back(chunk * 8);
            }
            const auto drain = [&](int s) {
                if (!pending_m[s]) return;
                LAPPER_CUDA_CHECK(cudaStreamSynchronize(sst[s].s));
                parallel_memcpy(counts_out + pending_lo[s], hout[s].p, pending_m[s] * 8);
                pending_m[s] = 0;
            };
            uint64_t done = 0;
            for (int i = 0; done < nq; i++) {
                const int s = i % kSlots;
                const uint64_t m = std::min(chunk, nq - done);
                drain(s);
                parallel_memcpy(hqs[s].p, q_start + done, m * 8);
                parallel_memcpy(hqe[s].p, q_stop + done, m * 8);
                LAPPER_CUDA_CHECK(cudaMemcpyAsync(dqs[s].p, hqs[s].p, m * 8, cudaMemcpyHostToDevice, sst[s].s));
                LAPPER_CUDA_CHECK(cudaMemcpyAsync(dqe[s].p, hqe[s].p, m * 8, cudaMemcpyHostToDevice, sst[s].s));
                const int rc = lapper64_count_batch_dev(l, dqs[s].p, dqe[s].p, m, dout[s].p, sst[s].s);
                if (rc != LAPPER_OK) throw Error(rc, last_error_slot());
                LAPPER_CUDA_CHECK(cudaMemcpyAsync(hout[s].p, dout[s].p, m * 8, cudaMemcpyDeviceToHost, sst[s].s));
                pending_lo[s] = done, pending_m[s] = m;
                done += m;
            }
            for (int k = 0; k < kSlots; k++) drain((int)((div_up(nq, chunk) + k) % kSlots));  // oldest chunk first
            for (int s = 0; s < used; s++) {
                dqs[s].release(), dqe[s].release(), dout[s].release();
                LAPPER_CUDA_CHECK(cudaStreamSynchronize(sst[s].s));
            }
            return;
        }
        OwnedStream st[1];  // slot 0; the other slots' streams only for batches of several chunks
        std::unique_ptr<OwnedStream> more[kSlots];
        cudaStream_t ss[kSlots] = {st[0].s, nullptr, nullptr};
        for (int s = 1; s < used; s++) more[s].reset(new OwnedStream()), ss[s] = more[s]->s;
        Scratch<uint64_t> dqs[kSlots], dqe[kSlots], dout[kSlots];
        for (int s = 0; s < used; s++) dqs[s].alloc(chunk, ss[s]), dqe[s].alloc(chunk, ss[s]), dout[s].alloc(chunk, ss[s]);
        uint64_t done = 0;
        for (int i = 0; done < nq; i++) {
            const int s = i % kSlots;
            const uint64_t m = std::min(chunk, nq - done);
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(dqs[s].p, q_start + done, m * 8, cudaMemcpyHostToDevice, ss[s]));
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(dqe[s].p, q_stop + done, m * 8, cudaMemcpyHostToDevice, ss[s]));
            // (a batch is not split between the two engines: the first chunk decides, see narrow_of)
            const int rc = lapper64_count_batch_dev(l, dqs[s].p, dqe[s].p, m, dout[s].p, ss[s]);
            if (rc != LAPPER_OK) throw Error(rc, last_error_slot());
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(counts_out + done, dout[s].p, m * 8, cudaMemcpyDeviceToHost, ss[s]));
            done += m;
        }
        for (int s = 0; s < used; s++) {
            dqs[s].release(), dqe[s].release(), dout[s].release();
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(ss[s]));
        }
    });
}

int lapper64_find_batch_dev(const lapper64_t *l, const uint64_t *d_q_start, const uint64_t *d_q_stop, uint64_t nq,
                            uint64_t *d_offsets, uint32_t *d_indices, uint64_t indices_capacity, uint64_t *total_out, void *stream) {
    return guarded_call([&] {
        check64(l);
        LAPPER_REQUIRE(d_offsets && total_out && (nq == 0 || (d_q_start && d_q_stop)), "find_batch_dev: null pointer");
        LAPPER_REQUIRE(d_indices || indices_capacity == 0, "find_batch_dev: null position buffer with a capacity");
        DeviceGuard dg(l->device);
        cudaStream_t s = (cudaStream_t)stream;
        *total_out = 0;
        if (lapper_t *h = narrow_of(l, nq, s)) {
            NarrowQueries q(l, d_q_start, d_q_stop, nq, s);
            const int rc = lapper_find_batch_u32_dev(h, q.a.p, q.b.p, nq, d_offsets, d_indices, indices_capacity, total_out, stream);
            if (rc != LAPPER_OK) throw Error(rc, last_error_slot());  // (LAPPER_ERR_CAPACITY included: same contract)
            return;
        }
        find_offsets64(l, d_q_start, d_q_stop, nq, d_offsets, s);
        if (d_indices) find_emit64(l, d_q_start, d_q_stop, nq, d_offsets, d_indices, indices_capacity, s);
        uint64_t total = 0;
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(&total, d_offsets + nq, 8, cudaMemcpyDeviceToHost, s));
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(s));
        *total_out = total;
        if (d_indices && total > indices_capacity)
            throw Error(LAPPER_ERR_CAPACITY, "find_batch_dev: " + std::to_string(total) + " positions do not fit the buffer of " +
                                                 std::to_string(indices_capacity) + " (offsets and *total_out are complete)");
    });
}

int lapper64_find_batch(const lapper64_t *l, const uint64_t *q_start, const uint64_t *q_stop, uint64_t nq, lapper_result_t **out) {
    return guarded_call([&] {
        check64(l);
        LAPPER_REQUIRE(out, "null result handle");
        *out = nullptr;
        LAPPER_REQUIRE(nq == 0 || (q_start && q_stop), "find_batch: null pointer");
        DeviceGuard dg(l->device);
        OwnedStream st;
        Scratch<uint64_t> dqs(nq, st.s), dqe(nq, st.s);
        if (nq) {
            copy_to_device(dqs.p, q_start, nq * 8, st.s);
            copy_to_device(dqe.p, q_stop, nq * 8, st.s);
        }
        if (lapper_t *h = narrow_of(l, nq, st.s)) {
            NarrowQueries q(l, dqs.p, dqe.p, nq, st.s);
            dqs.release(), dqe.release();
            const int rc = lapper_find_batch_u32_devq(h, q.a.p, q.b.p, nq, st.s, out);
            if (rc != LAPPER_OK) throw Error(rc, last_error_slot());
            return;
        }
        std::unique_ptr<lapper_result_t> r(new lapper_result_t());
        r->device = l->device;
        r->nq = nq;
        uint64_t bytes = 0;
        try {
            r->d_offsets = pool_alloc<uint64_t>(nq + 1, bytes, st.s);
            find_offsets64(l, dqs.p, dqe.p, nq, r->d_offsets, st.s);
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(&r->total, r->d_offsets + nq, 8, cudaMemcpyDeviceToHost, st.s));
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
            r->d_indices = pool_alloc<uint32_t>(r->total, bytes, st.s);
            find_emit64(l, dqs.p, dqe.p, nq, r->d_offsets, r->d_indices, r->total, st.s);
            dqs.release(), dqe.release();
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
        } catch (...) {
            pool_free(r->d_offsets, st.s), pool_free(r->d_indices, st.s);
            cudaStreamSynchronize(st.s);
            throw;
        }
        *out = r.release();
    });
}

int lapper64_cov(const lapper64_t *l, uint64_t *cov_out) {
    return guarded_call([&] {
        check64(l);
        LAPPER_REQUIRE(cov_out, "null pointer");
        if (l->cov_is_some) {  // src/lib.rs:311-316
            *cov_out = l->cov;
            return;
        }
        need_valid(l, "cov");
        DeviceGuard dg(l->device);
        OwnedStream st;
        *cov_out = cov_of(l, st.s);
    });
}

int lapper64_set_cov(lapper64_t *l, uint64_t *cov_out) {
    return guarded_call([&] {
        check64(l);
        need_valid(l, "set_cov");
        DeviceGuard dg(l->device);
        OwnedStream st;
        l->cov = cov_of(l, st.s);  // src/lib.rs:320-324
        l->cov_is_some = true;
        if (cov_out) *cov_out = l->cov;
    });
}

/* the cached fields of src/lib.rs:119-122 (`cov: Option<I>`, `overlaps_merged`), for (de)serialisation */
int lapper64_get_cached_state(const lapper64_t *l, int *cov_is_some_out, uint64_t *cov_out, int *overlaps_merged_out) {
    return guarded_call([&] {
        check64(l);
        if (cov_is_some_out) *cov_is_some_out = l->cov_is_some ? 1 : 0;
        if (cov_out) *cov_out = l->cov_is_some ? l->cov : 0u;
        if (overlaps_merged_out) *overlaps_merged_out = l->overlaps_merged ? 1 : 0;
    });
}

int lapper64_restore_cached_state(lapper64_t *l, int cov_is_some, uint64_t cov, int overlaps_merged) {
    return guarded_call([&] {
        check64(l);
        l->cov_is_some = cov_is_some != 0;
        l->cov = cov_is_some ? cov : 0u;
        l->overlaps_merged = overlaps_merged != 0;
    });
}

int lapper64_merge_overlaps(lapper64_t *l) {
    return guarded_call([&] {
        check64(l);
        if (l->n == 0) return;  // src/lib.rs:366: overlaps_merged is set only when there is a first interval
        DeviceGuard dg(l->device);
        OwnedStream st;
        const bool inverted = l->has_inverted;  // (merged_of runs the reference's loop literally for such sets)
        Merged64 mg(st.s);
        merged_of(l, mg, st.s);
        // the new index is completed aside and swapped in only when nothing can fail any more
        lapper64 next;
        next.device = l->device;
        next.n = mg.m;
        next.S = mg.s, next.Eiv = mg.e, next.perm = mg.perm;
        mg.s = mg.e = nullptr, mg.perm = nullptr;
        try {
            uint64_t bytes = 0;
            // the merged runs are disjoint and ascending: their stops are sorted, and the prefix max is the stops themselves
            next.E = pool_alloc<uint64_t>(next.n, bytes, st.s);
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(next.E, next.Eiv, next.n * 8, cudaMemcpyDeviceToDevice, st.s));
            if (inverted && next.n > 1) {  // a run may have kept an inverted interval's stop: `stops` has to be sorted (:396-406)
                Scratch<uint64_t> tmp(next.n, st.s);
                launch_sort_u64(next.E, tmp.p, next.n, st.s);
            }
            finish_from_sorted(&next, st.s);
        } catch (...) {
            free_arrays(&next, st.s);
            cudaStreamSynchronize(st.s);
            throw;
        }
        drop_narrow(l);
        free_arrays(l, st.s);
        l->n = next.n;
        l->S = next.S, l->Eiv = next.Eiv, l->E = next.E, l->P = next.P, l->perm = next.perm;
        l->max_len = next.max_len, l->has_inverted = next.has_inverted, l->max_coord = next.max_coord, l->min_coord = next.min_coord;
        l->overlaps_merged = true;
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
    });
}

int lapper64_union_and_intersect(const lapper64_t *a, const lapper64_t *b, uint64_t *union_out, uint64_t *intersect_out) {
    return guarded_call([&] {
        check64(a), check64(b);
        LAPPER_REQUIRE(a->device == b->device, "union_and_intersect: both lappers must live on the same device");
        need_valid(a, "union_and_intersect"), need_valid(b, "union_and_intersect");
        DeviceGuard dg(a->device);
        OwnedStream st;
        uint64_t x = 0;
        if (a->n && b->n) {
            Merged64 ma(st.s), mb(st.s);
            merged_of(a, ma, st.s), merged_of(b, mb, st.s);
            Scratch<uint64_t> cum(mb.m, st.s);
            Scratch<unsigned long long> d(1, st.s);
            scan64<Add64>(RunLen64{mb.s, mb.e}, mb.m, StoreExcl{cum.p}, nullptr, st.s);
            LAPPER_CUDA_CHECK(cudaMemsetAsync(d.p, 0, 8, st.s));
            intersect64_kernel<<<capped(ma.m), kT, 0, st.s>>>(ma.s, ma.e, ma.m, mb.s, mb.e, cum.p, mb.m, d.p);
            LAPPER_CUDA_CHECK(cudaGetLastError());
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(&x, d.p, 8, cudaMemcpyDeviceToHost, st.s));
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
        }
        const uint64_t ca = a->cov_is_some ? a->cov : cov_of(a, st.s);
        const uint64_t cb = b->cov_is_some ? b->cov : cov_of(b, st.s);
        if (union_out) *union_out = ca + cb - x;  // src/lib.rs:532 / :542
        if (intersect_out) *intersect_out = x;
    });
}

int lapper64_insert(lapper64_t *l, uint64_t start, uint64_t stop, uint64_t *input_position_out) {
    return guarded_call([&] {
        check64(l);
        DeviceGuard dg(l->device);
        OwnedStream st;
        const uint64_t n = l->n;
        LAPPER_REQUIRE(n + 1 < 0xFFFFFFF0ull && l->next_input_id < 0xFFFFFFF0ull, "lapper64_insert: too many intervals");
        uint64_t pos[2] = {0, 0};
        if (n) {
            Scratch<uint64_t> d_pos(2, st.s);
            insert_positions64_kernel<<<1, 1, 0, st.s>>>(l->S, l->Eiv, l->E, n, start, stop, d_pos.p);
            LAPPER_CUDA_CHECK(cudaGetLastError());
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(pos, d_pos.p, 16, cudaMemcpyDeviceToHost, st.s));
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
        }
        const uint64_t id = l->next_input_id;
        // the new index is completed aside: a failure leaves the handle as it was
        lapper64 next;
        next.device = l->device;
        next.n = n + 1;
        try {
            uint64_t bytes = 0;
            next.S = pool_alloc<uint64_t>(n + 1, bytes, st.s);
            next.Eiv = pool_alloc<uint64_t>(n + 1, bytes, st.s);
            next.E = pool_alloc<uint64_t>(n + 1, bytes, st.s);
            next.perm = pool_alloc<uint32_t>(n + 1, bytes, st.s);
            const uint32_t g = capped(n + 1);
            insert_into64_kernel<uint64_t><<<g, kT, 0, st.s>>>(next.S, l->S, n, pos[0], start);
            insert_into64_kernel<uint64_t><<<g, kT, 0, st.s>>>(next.Eiv, l->Eiv, n, pos[0], stop);
            insert_into64_kernel<uint64_t><<<g, kT, 0, st.s>>>(next.E, l->E, n, pos[1], stop);
            insert_into64_kernel<uint32_t><<<g, kT, 0, st.s>>>(next.perm, l->perm, n, pos[0], (uint32_t)id);
            LAPPER_CUDA_CHECK(cudaGetLastError());
            finish_from_sorted(&next, st.s);  // max_len (src/lib.rs:264-267) falls out of it
        } catch (...) {
            free_arrays(&next, st.s);
            cudaStreamSynchronize(st.s);
            throw;
        }
        drop_narrow(l);
        free_arrays(l, st.s);
        l->n = next.n;
        l->S = next.S, l->Eiv = next.Eiv, l->E = next.E, l->P = next.P, l->perm = next.perm;
        l->max_len = next.max_len, l->has_inverted = next.has_inverted, l->max_coord = next.max_coord, l->min_coord = next.min_coord;
        l->next_input_id = id + 1;
        l->cov_is_some = false;      // src/lib.rs:271
        l->overlaps_merged = false;  // src/lib.rs:272
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
        if (input_position_out) *input_position_out = id;
    });
}

int lapper64_seek_cursor(const lapper64_t *l, uint64_t start, uint64_t *cursor) {
    return guarded_call([&] {
        check64(l);
        LAPPER_REQUIRE(cursor, "null cursor");
        if (l->n == 0) return;  // lower_bound over an empty vector is 0 and the while-loop never runs (src/lib.rs:658-671)
        DeviceGuard dg(l->device);
        OwnedStream st;
        Scratch<uint64_t> d_c(1, st.s);
        seek_cursor64_kernel<<<1, 1, 0, st.s>>>(l->S, l->n, l->max_len, start, *cursor, d_c.p);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(cursor, d_c.p, 8, cudaMemcpyDeviceToHost, st.s));
        d_c.release();
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(st.s));
    });
}

int lapper64_depth(const lapper64_t *l, uint64_t *n_out, uint64_t **starts_out, uint64_t **stops_out, uint32_t **depth_out) {
    return guarded_call([&] {
        check64(l);
        LAPPER_REQUIRE(n_out && starts_out && stops_out && depth_out, "lapper64_depth: null pointer");
        *n_out = 0;
        *starts_out = *stops_out = nullptr;
        *depth_out = nullptr;
        LAPPER_REQUIRE(l->n > 0, "depth() on an empty lapper: the reference panics here (src/lib.rs:744)");
        need_valid(l, "depth");
        DeviceGuard dg(l->device);
        OwnedStream st;
        cudaStream_t s = st.s;
        const uint64_t n = l->n, n2 = 2 * n;
        Scratch<uint64_t> X(n2, s), tmp(n2, s), U(n2, s), d_count(1, s);
        Scratch<uint32_t> D(n2, s), HS(n2, s);
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(X.p, l->S, n * 8, cudaMemcpyDeviceToDevice, s));
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(X.p + n, l->E, n * 8, cudaMemcpyDeviceToDevice, s));
        launch_sort_u64(X.p, tmp.p, n2, s);
        uint64_t m = 0, ne = 0, nr = 0;
        scan64<Add64>(Unique64{X.p}, n2, UniqueSink64{X.p, U.p}, d_count.p, s);
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(&m, d_count.p, 8, cudaMemcpyDeviceToHost, s));
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(s));
        depth_eval64_kernel<<<capped(m), kT, 0, s>>>(l->S, l->E, n, U.p, m, D.p, HS.p);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        Scratch<uint64_t> eu(m + 1, s);
        Scratch<uint32_t> ed(m + 1, s), ez(m + 1, s);
        const EmitPoint64 epf{D.p, HS.p};
        scan64<Add64>(epf, m, EmitPointSink64{U.p, epf, eu.p, ed.p, ez.p}, d_count.p, s);
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(&ne, d_count.p, 8, cudaMemcpyDeviceToHost, s));
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(s));
        // the runs: counted first (the outputs are sized from the count), then written
        const Run64 rf{ed.p, ez.p};
        const uint64_t ntiles = div_up(ne ? ne : 1, kTile64);
        Scratch<uint64_t> agg(ntiles, s);
        if (ne) {
            tile_reduce64_kernel<Add64, Run64><<<(uint32_t)ntiles, kT, 0, s>>>(rf, ne, agg.p);
            tile_scan64_kernel<Add64><<<1, kT, 0, s>>>(agg.p, ntiles, d_count.p);
            LAPPER_CUDA_CHECK(cudaGetLastError());
            LAPPER_CUDA_CHECK(cudaMemcpyAsync(&nr, d_count.p, 8, cudaMemcpyDeviceToHost, s));
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(s));
        }
        Scratch<uint64_t> os(nr, s), oe(nr, s);
        Scratch<uint32_t> od(nr, s);
        if (nr) {
            tile_apply64_kernel<Add64, Run64, RunSink64><<<(uint32_t)ntiles, kT, 0, s>>>(rf, ne, agg.p,
                                                                                        RunSink64{eu.p, ed.p, ez.p, os.p, oe.p, od.p});
            LAPPER_CUDA_CHECK(cudaGetLastError());
        }
        uint64_t *hs = (uint64_t *)malloc(std::max<uint64_t>(nr, 1) * 8), *he = (uint64_t *)malloc(std::max<uint64_t>(nr, 1) * 8);
        uint32_t *hd = (uint32_t *)malloc(std::max<uint64_t>(nr, 1) * 4);
        try {
            if (!hs || !he || !hd) throw std::bad_alloc();
            if (nr) {
                LAPPER_CUDA_CHECK(cudaMemcpyAsync(hs, os.p, nr * 8, cudaMemcpyDeviceToHost, s));
                LAPPER_CUDA_CHECK(cudaMemcpyAsync(he, oe.p, nr * 8, cudaMemcpyDeviceToHost, s));
                LAPPER_CUDA_CHECK(cudaMemcpyAsync(hd, od.p, nr * 4, cudaMemcpyDeviceToHost, s));
            }
            LAPPER_CUDA_CHECK(cudaStreamSynchronize(s));
        } catch (...) {
            free(hs), free(he), free(hd);
            throw;
        }
        *n_out = nr;
        *starts_out = hs, *stops_out = he, *depth_out = hd;  // freed with lapper_host_free
    });
}

/* lapper_tune_query_length for a u64 handle: takes effect where the u32 engine answers (narrowed sets) */
int lapper64_tune_query_length(lapper64_t *l, uint32_t typical_query_len) {
    return guarded_call([&] {
        check64(l);
        l->query_len_hint = typical_query_len;
        if (lapper_t *h = l->narrow.load(std::memory_order_acquire)) {
            const int rc = lapper_tune_query_length(h, typical_query_len);
            if (rc != LAPPER_OK) throw Error(rc, last_error_slot());
        }
    });
}

/* which engine answers this handle's batches: 1 = the u32 engine is built (coordinates fit 32 bits), 0 = the 64-bit kernels */
int lapper64_is_narrow(const lapper64_t *l) { return l && l->narrow.load(std::memory_order_acquire) != nullptr; }

/* tests: 0 = never narrow, 1 = from the first batch of 4096 queries on (default), 2 = always; returns the old mode */
int lapper64_debug_set_narrow_mode(int mode) { return g_narrow_mode.exchange(std::min(std::max(mode, 0), 2)); }

}  // extern "C"
