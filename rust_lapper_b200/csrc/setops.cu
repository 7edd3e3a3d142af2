// setops.cu -- merge_overlaps, cov and union_and_intersect as scans over the sorted index.
//
// Closed forms (each verified against the literal loops by
// tests/test_oracle_golden.py::test_merge_cov_union_closed_forms); they need start <= stop:
//   P_i     = max(stop_0 .. stop_i)                       inclusive prefix max
//   head(i) = i == 0 || P_{i-1} < start_i                  strict: touching intervals merge (src/lib.rs:370)
//   merged run r = (start_head, P_last-of-run), val of the head (src/lib.rs:363-391)
//   cov     = sum over runs of (P_last - start_head)       (src/lib.rs:327-350; u32, wrapping)
//   A n B   = sum over merged runs a of F_B(a.stop) - F_B(a.start), F_B(x) = |union(B) n [0, x)|
//             (what src/lib.rs:517-543 computes by enumerating pairs, in either branch)
#include <algorithm>

#include "index.cuh"
#include "merge_sequential.cuh"

namespace lap {

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;

struct MaxOp {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return max(a, b); }
    static __device__ __forceinline__ uint32_t identity() { return 0u; }
};
struct AddOp {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a + b; }
    static __device__ __forceinline__ uint32_t identity() { return 0u; }
};

// input functors: element i of the sequence being scanned
struct LoadArray {
    const uint32_t *p;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return p[i]; }
};
struct HeadFlag {  // 1 where a merged run starts
    const uint32_t *S;
    const uint32_t *P;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return (i == 0 || P[i - 1] < S[i]) ? 1u : 0u; }
};
struct RunLength {  // lengths of already-merged runs
    const uint32_t *s;
    const uint32_t *e;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return e[i] - s[i]; }
};

// length bin of an interval: 0 for length 0 (or inverted, checked_sub -> 0), else its bit length
__device__ __forceinline__ uint32_t len_bin(uint32_t s, uint32_t e) { return e > s ? 32u - __clz(e - s) : 0u; }

struct ClassTable {
    uint8_t of_bin[36];
};
struct ClassFlag {  // 1 where the interval belongs to class c
    const uint32_t *S;
    const uint32_t *E;
    ClassTable t;
    uint32_t c;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return t.of_bin[len_bin(S[i], E[i])] == c ? 1u : 0u; }
};
struct ClassSink {
    const uint32_t *S;
    const uint32_t *E;
    uint32_t off;
    uint32_t *cs, *ce, *cg;
    __device__ __forceinline__ void operator()(uint64_t i, uint32_t flag, uint32_t inclusive) const {
        if (flag) {
            const uint32_t pos = off + inclusive - 1u;
            cs[pos] = S[i];
            ce[pos] = E[i];
            cg[pos] = (uint32_t)i;
        }
    }
};

template <typename Op>
__device__ __forceinline__ uint32_t block_scan_inclusive(uint32_t x, Op op, uint32_t *warp_tot, uint32_t &block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc = op(inc, y);
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    uint32_t prefix = Op::identity(), total = Op::identity();
#pragma unroll
    for (int i = 0; i < kThreads / 32; i++) {
        const uint32_t t = warp_tot[i];
        if (i < w) prefix = op(prefix, t);
        total = op(total, t);
    }
    block_total = total;
    __syncthreads();
    return op(prefix, inc);
}

// phase 1: per-tile aggregate
template <typename Op, typename In>
__global__ void __launch_bounds__(kThreads) tile_reduce_kernel(In in, uint64_t n, uint32_t *__restrict__ tile_agg) {
    __shared__ uint32_t warp_tot[kThreads / 32];
    Op op;
    const uint64_t base = (uint64_t)blockIdx.x * kTile + (uint64_t)threadIdx.x * kItems;
    uint32_t acc = Op::identity();
#pragma unroll
    for (int k = 0; k < kItems; k++)
        if (base + k < n) acc = op(acc, in(base + k));
    uint32_t total;
    block_scan_inclusive(acc, op, warp_tot, total);
    if (threadIdx.x == 0) tile_agg[blockIdx.x] = total;
}

// phase 2: exclusive scan of the tile aggregates (single block), grand total to *total_out
template <typename Op>
__global__ void __launch_bounds__(kThreads) tile_scan_kernel(uint32_t *__restrict__ tile_agg, uint64_t ntiles,
                                                             uint32_t *__restrict__ total_out) {
    __shared__ uint32_t warp_tot[kThreads / 32];
    __shared__ uint32_t carry_s;
    Op op;
    if (threadIdx.x == 0) carry_s = Op::identity();
    __syncthreads();
    for (uint64_t base = 0; base < ntiles; base += kThreads) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t x = i < ntiles ? tile_agg[i] : Op::identity();
        uint32_t total;
        const uint32_t inc = block_scan_inclusive(x, op, warp_tot, total);
        const uint32_t carry = carry_s;
        // exclusive prefix = carry (+) everything before me in this chunk
        const uint32_t before = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
        uint32_t excl;
        if ((threadIdx.x & 31) == 0) {
            // first lane of a warp: rebuild from the warp totals
            uint32_t p = Op::identity();
            for (int wi = 0; wi < (int)(threadIdx.x >> 5); wi++) p = op(p, warp_tot[wi]);
            excl = op(carry, p);
        } else {
            excl = op(carry, before);
        }
        if (i < ntiles) tile_agg[i] = excl;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = op(carry, total);
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

// phase 3: rescan each tile with its carry-in and hand every element's inclusive value to `sink`
template <typename Op, typename In, typename Sink>
__global__ void __launch_bounds__(kThreads) tile_apply_kernel(In in, uint64_t n, const uint32_t *__restrict__ tile_excl,
                                                              Sink sink) {
    __shared__ uint32_t warp_tot[kThreads / 32];
    Op op;
    const uint64_t base = (uint64_t)blockIdx.x * kTile + (uint64_t)threadIdx.x * kItems;
    uint32_t v[kItems];
    uint32_t acc = Op::identity();
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        v[k] = base + k < n ? in(base + k) : Op::identity();
        acc = op(acc, v[k]);
    }
    uint32_t total;
    const uint32_t inc = block_scan_inclusive(acc, op, warp_tot, total);
    // exclusive prefix of this thread = carry-in (+) inclusive of the previous thread
    uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if ((threadIdx.x & 31) == 0) {
        prev = Op::identity();
        for (int wi = 0; wi < (int)(threadIdx.x >> 5); wi++) prev = op(prev, warp_tot[wi]);
    }
    uint32_t run = op(tile_excl[blockIdx.x], prev);
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        run = op(run, v[k]);
        if (base + k < n) sink(base + k, v[k], run);
    }
}

struct StoreSink {
    uint32_t *out;
    __device__ __forceinline__ void operator()(uint64_t i, uint32_t, uint32_t inclusive) const { out[i] = inclusive; }
};
struct StoreExclusiveSink {
    uint32_t *out;
    __device__ __forceinline__ void operator()(uint64_t i, uint32_t x, uint32_t inclusive) const { out[i] = inclusive - x; }
};
// run compaction: element i belongs to run (inclusive head count - 1)
struct MergeSink {
    const uint32_t *S, *P, *perm;
    uint64_t n;
    uint32_t *m_starts, *m_stops, *m_perm;
    __device__ __forceinline__ void operator()(uint64_t i, uint32_t is_head, uint32_t heads_inclusive) const {
        const uint32_t r = heads_inclusive - 1u;
        if (is_head) {
            m_starts[r] = S[i];
            m_perm[r] = perm[i];
        }
        const bool last = (i + 1 == n) || (P[i] < S[i + 1]);
        if (last) m_stops[r] = P[i];
    }
};

// ---- depth(): functors of the two compactions (see compute_depth)
struct UniqueFlag {  // first occurrence of each coordinate in the sorted event list
    const uint32_t *X;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return (i == 0 || X[i] != X[i - 1]) ? 1u : 0u; }
};
struct UniqueSink {
    const uint32_t *X;
    uint32_t *U;
    __device__ __forceinline__ void operator()(uint64_t i, uint32_t flag, uint32_t inclusive) const {
        if (flag) U[inclusive - 1] = X[i];
    }
};
// emit points among the unique coordinates: depth changes, and zero-length merged runs (depth 0 before and
// at the coordinate, with an interval starting there) other than the very first run
struct EmitPointFlag {
    const uint32_t *D;
    const uint32_t *has_start;
    __device__ __forceinline__ bool zero_run(uint64_t t) const { return t > 0 && D[t] == 0 && D[t - 1] == 0 && has_start[t]; }
    __device__ __forceinline__ uint32_t operator()(uint64_t t) const {
        return (t == 0 || D[t] != D[t - 1] || zero_run(t)) ? 1u : 0u;
    }
};
struct EmitPointSink {
    const uint32_t *U;
    EmitPointFlag f;
    uint32_t *eu, *ed, *ez;
    __device__ __forceinline__ void operator()(uint64_t t, uint32_t flag, uint32_t inclusive) const {
        if (flag) {
            eu[inclusive - 1] = U[t];
            ed[inclusive - 1] = f.D[t];
            ez[inclusive - 1] = f.zero_run(t) ? 1u : 0u;
        }
    }
};
struct RunFlag {  // emit points that produce a run: positive depth, or a zero-length merged run
    const uint32_t *ed, *ez;
    __device__ __forceinline__ uint32_t operator()(uint64_t q) const { return (ed[q] > 0 || ez[q]) ? 1u : 0u; }
};
struct RunSink {
    const uint32_t *eu, *ed, *ez;
    uint32_t *os, *oe, *od;
    __device__ __forceinline__ void operator()(uint64_t q, uint32_t flag, uint32_t inclusive) const {
        if (flag) {
            const uint32_t r = inclusive - 1;
            os[r] = eu[q];
            // a positive-depth run ends at the next emit point, which is always a depth change
            oe[r] = ez[q] ? eu[q] : eu[q + 1];
            od[r] = ed[q];
        }
    }
};

// depth at every unique coordinate u: #{start <= u} - #{stop <= u}, and whether an interval starts at u
__global__ void __launch_bounds__(kThreads) depth_eval_kernel(const uint32_t *__restrict__ S, const uint32_t *__restrict__ E,
                                                              uint32_t n, const uint32_t *__restrict__ U, uint64_t m,
                                                              uint32_t *__restrict__ D, uint32_t *__restrict__ has_start) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t u = U[t];
        const uint32_t ub_s = upper_bound_u32(S, 0, n, u);
        D[t] = ub_s - upper_bound_u32(E, 0, n, u);
        has_start[t] = lower_bound_u32(S, 0, n, u) < ub_s ? 1u : 0u;
    }
}

template <typename Op, typename In, typename Sink>
void scan3(In in, uint64_t n, Sink sink, uint32_t *d_total, cudaStream_t stream) {
    if (n == 0) {
        if (d_total) LAPPER_CUDA_CHECK(cudaMemsetAsync(d_total, 0, 4, stream));
        return;
    }
    const uint64_t ntiles = div_up(n, kTile);
    Scratch<uint32_t> agg(ntiles, stream);
    tile_reduce_kernel<Op, In><<<(uint32_t)ntiles, kThreads, 0, stream>>>(in, n, agg.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    tile_scan_kernel<Op><<<1, kThreads, 0, stream>>>(agg.p, ntiles, d_total);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    tile_apply_kernel<Op, In, Sink><<<(uint32_t)ntiles, kThreads, 0, stream>>>(in, n, agg.p, sink);
    LAPPER_CUDA_CHECK(cudaGetLastError());
}

// cov = sum_{last of run} P_i - sum_{head} S_i  (wrapping u32)
// (over positions [lo, hi) of the n intervals: the prefix max is global, so a chunk's partial sum needs nothing from
// the other chunks -- the multi-GPU form of cov is the sum of the replicas' chunk sums)
__global__ void __launch_bounds__(kThreads) cov_kernel(const uint32_t *__restrict__ S, const uint32_t *__restrict__ P,
                                                       uint64_t n, uint64_t lo, uint64_t hi, uint32_t *__restrict__ out) {
    uint32_t acc = 0;
    for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = S[i], p = P[i];
        const bool head = (i == 0) || (P[i - 1] < s);
        const bool last = (i + 1 == n) || (p < S[i + 1]);
        if (head) acc -= s;
        if (last) acc += p;
    }
    acc = __reduce_add_sync(0xFFFFFFFFu, acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// covered length of union(B) below x; B given as disjoint sorted runs with exclusive prefix lengths
__device__ __forceinline__ uint32_t covered_below(const uint32_t *__restrict__ bs, const uint32_t *__restrict__ be,
                                                  const uint32_t *__restrict__ cum, uint32_t m, uint32_t x) {
    const uint32_t j = lower_bound_u32(bs, 0, m, x);  // runs starting below x
    if (j == 0) return 0;
    const uint32_t s = bs[j - 1], e = be[j - 1];
    return cum[j - 1] + (min(x, e) - s);
}

__global__ void __launch_bounds__(kThreads) intersect_kernel(const uint32_t *__restrict__ as, const uint32_t *__restrict__ ae,
                                                             uint32_t ma, const uint32_t *__restrict__ bs,
                                                             const uint32_t *__restrict__ be,
                                                             const uint32_t *__restrict__ cum, uint32_t mb,
                                                             uint32_t *__restrict__ out) {
    uint32_t acc = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < ma; r += gridDim.x * blockDim.x)
        acc += covered_below(bs, be, cum, mb, ae[r]) - covered_below(bs, be, cum, mb, as[r]);
    acc = __reduce_add_sync(0xFFFFFFFFu, acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

uint32_t reduce_grid(uint64_t n) {
    uint64_t g = div_up(n ? n : 1, kThreads);
    const uint64_t cap = (uint64_t)sm_count() * 8;
    return (uint32_t)(g < cap ? g : cap);
}

struct DevArrays {  // three pool arrays, returned to the pool (stream-ordered) unless handed on
    uint32_t *s = nullptr, *e = nullptr, *perm = nullptr;
    cudaStream_t stream;
    explicit DevArrays(cudaStream_t st) : stream(st) {}
    void alloc(uint64_t count) {
        uint64_t bytes = 0;
        s = pool_alloc<uint32_t>(count, bytes, stream);
        e = pool_alloc<uint32_t>(count, bytes, stream);
        perm = pool_alloc<uint32_t>(count, bytes, stream);
    }
    ~DevArrays() {
        pool_free(s, stream);
        pool_free(e, stream);
        pool_free(perm, stream);
    }
};

}  // namespace

void launch_class_partition(DeviceIndex &ix, const uint8_t *class_of_bin, cudaStream_t stream) {
    ClassTable t;
    for (int b = 0; b < 36; b++) t.of_bin[b] = b < 33 ? class_of_bin[b] : 0;
    for (uint32_t c = 0; c < ix.ncls; c++) {
        if (ix.cls_n[c] == 0) continue;
        scan3<AddOp>(ClassFlag{ix.starts, ix.stops_by_iv, t, c}, ix.n,
                     ClassSink{ix.starts, ix.stops_by_iv, ix.cls_off[c], ix.cls_starts, ix.cls_stops, ix.cls_gidx}, nullptr,
                     stream);
    }
}

void launch_exclusive_sum(const uint32_t *in, uint32_t *out, uint64_t n, cudaStream_t stream) {
    scan3<AddOp>(LoadArray{in}, n, StoreExclusiveSink{out}, nullptr, stream);
}

void launch_prefix_max(const uint32_t *in, uint32_t *out, uint64_t n, cudaStream_t stream) {
    scan3<MaxOp>(LoadArray{in}, n, StoreSink{out}, nullptr, stream);
}

uint32_t compute_cov(const DeviceIndex &ix, cudaStream_t stream) { return compute_cov_range(ix, 0, ix.n, stream); }

uint32_t compute_cov_range(const DeviceIndex &ix, uint64_t lo, uint64_t hi, cudaStream_t stream) {
    hi = std::min<uint64_t>(hi, ix.n);
    if (ix.n == 0 || lo >= hi) return 0;
    Scratch<uint32_t> d_out(1, stream);
    LAPPER_CUDA_CHECK(cudaMemsetAsync(d_out.p, 0, 4, stream));
    cov_kernel<<<reduce_grid(hi - lo), kThreads, 0, stream>>>(ix.starts, ix.prefix_max, ix.n, lo, hi, d_out.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    uint32_t cov = 0;
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(&cov, d_out.p, 4, cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    return cov;
}

uint64_t compute_merged(const DeviceIndex &ix, uint32_t **m_starts, uint32_t **m_stops, uint32_t **m_perm,
                        cudaStream_t stream) {
    *m_starts = *m_stops = *m_perm = nullptr;
    if (ix.n == 0) return 0;
    if (ix.has_inverted) {
        // some stop < start: the reference's own loop, one warp (merge_sequential.cuh)
        Scratch<unsigned long long> d_m(1, stream);
        merge_sequential_kernel<uint32_t, false><<<1, 32, 0, stream>>>(ix.starts, ix.stops_by_iv, ix.perm, ix.n, nullptr, nullptr,
                                                                      nullptr, d_m.p);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        unsigned long long m = 0;
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(&m, d_m.p, 8, cudaMemcpyDeviceToHost, stream));
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
        DevArrays out(stream);
        out.alloc(m);
        merge_sequential_kernel<uint32_t, true><<<1, 32, 0, stream>>>(ix.starts, ix.stops_by_iv, ix.perm, ix.n, out.s, out.e, out.perm,
                                                                     nullptr);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        *m_starts = out.s;
        *m_stops = out.e;
        *m_perm = out.perm;
        out.s = out.e = out.perm = nullptr;
        return m;
    }
    // the run count is needed before the outputs can be sized: count heads first
    Scratch<uint32_t> d_total(1, stream);
    const uint64_t n = ix.n;
    const uint64_t ntiles = div_up(n, kTile);
    Scratch<uint32_t> agg(ntiles, stream);
    HeadFlag flags{ix.starts, ix.prefix_max};
    tile_reduce_kernel<AddOp, HeadFlag><<<(uint32_t)ntiles, kThreads, 0, stream>>>(flags, n, agg.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    tile_scan_kernel<AddOp><<<1, kThreads, 0, stream>>>(agg.p, ntiles, d_total.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    uint32_t m = 0;
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(&m, d_total.p, 4, cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    DevArrays out(stream);
    out.alloc(m);
    MergeSink sink{ix.starts, ix.prefix_max, ix.perm, n, out.s, out.e, out.perm};
    tile_apply_kernel<AddOp, HeadFlag, MergeSink><<<(uint32_t)ntiles, kThreads, 0, stream>>>(flags, n, agg.p, sink);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    *m_starts = out.s;
    *m_stops = out.e;
    *m_perm = out.perm;
    out.s = out.e = out.perm = nullptr;
    return m;
}

// Closed form of IterDepth (src/lib.rs:735-784), verified against the literal loop in
// tests/test_oracle_golden.py::test_depth_closed_form: over the sorted unique event coordinates u_t with
// depth D_t = #{start <= u_t} - #{stop <= u_t}, a run starts wherever D changes to a positive value and
// ends at the next change; a zero-length merged run (an isolated coordinate where only zero-length
// intervals sit) is reported as (u, u, 0) unless it is the first run of the set.
uint64_t compute_depth(const DeviceIndex &ix, uint32_t **d_starts, uint32_t **d_stops, uint32_t **d_depth,
                       cudaStream_t stream) {
    *d_starts = *d_stops = *d_depth = nullptr;
    const uint64_t n = ix.n, n2 = 2 * n;
    Scratch<uint32_t> X(n2, stream), tmp(n2, stream), U(n2, stream), D(n2, stream), HS(n2, stream);
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(X.p, ix.starts, n * 4, cudaMemcpyDeviceToDevice, stream));
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(X.p + n, ix.stops_sorted, n * 4, cudaMemcpyDeviceToDevice, stream));
    launch_sort_u32(X.p, tmp.p, n2, stream);
    Scratch<uint32_t> d_count(1, stream);
    uint32_t m = 0, ne = 0, nr = 0;
    scan3<AddOp>(UniqueFlag{X.p}, n2, UniqueSink{X.p, U.p}, d_count.p, stream);
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(&m, d_count.p, 4, cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    depth_eval_kernel<<<reduce_grid(m), kThreads, 0, stream>>>(ix.starts, ix.stops_sorted, (uint32_t)n, U.p, m, D.p, HS.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    Scratch<uint32_t> eu((uint64_t)m + 1, stream), ed((uint64_t)m + 1, stream), ez((uint64_t)m + 1, stream);
    const EmitPointFlag epf{D.p, HS.p};
    scan3<AddOp>(epf, m, EmitPointSink{U.p, epf, eu.p, ed.p, ez.p}, d_count.p, stream);
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(&ne, d_count.p, 4, cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    // count the runs first, then size the outputs
    const RunFlag rf{ed.p, ez.p};
    {
        const uint64_t ntiles = div_up(ne, kTile);
        Scratch<uint32_t> agg(ntiles, stream);
        tile_reduce_kernel<AddOp, RunFlag><<<(uint32_t)ntiles, kThreads, 0, stream>>>(rf, ne, agg.p);
        tile_scan_kernel<AddOp><<<1, kThreads, 0, stream>>>(agg.p, ntiles, d_count.p);
        LAPPER_CUDA_CHECK(cudaGetLastError());
        LAPPER_CUDA_CHECK(cudaMemcpyAsync(&nr, d_count.p, 4, cudaMemcpyDeviceToHost, stream));
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
        DevArrays out(stream);
        out.alloc(nr);
        tile_apply_kernel<AddOp, RunFlag, RunSink><<<(uint32_t)ntiles, kThreads, 0, stream>>>(
            rf, ne, agg.p, RunSink{eu.p, ed.p, ez.p, out.s, out.e, out.perm});
        LAPPER_CUDA_CHECK(cudaGetLastError());
        LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
        *d_starts = out.s;
        *d_stops = out.e;
        *d_depth = out.perm;
        out.s = out.e = out.perm = nullptr;
    }
    return nr;
}

uint32_t compute_intersect_measure(const DeviceIndex &a, const DeviceIndex &b, cudaStream_t stream) {
    if (a.n == 0 || b.n == 0) return 0;
    DevArrays ma(stream), mb(stream);
    const uint64_t na = compute_merged(a, &ma.s, &ma.e, &ma.perm, stream);
    const uint64_t nb = compute_merged(b, &mb.s, &mb.e, &mb.perm, stream);
    Scratch<uint32_t> cum(nb, stream), d_out(1, stream);
    scan3<AddOp>(RunLength{mb.s, mb.e}, nb, StoreExclusiveSink{cum.p}, nullptr, stream);
    LAPPER_CUDA_CHECK(cudaMemsetAsync(d_out.p, 0, 4, stream));
    intersect_kernel<<<reduce_grid(na), kThreads, 0, stream>>>(ma.s, ma.e, (uint32_t)na, mb.s, mb.e, cum.p,
                                                               (uint32_t)nb, d_out.p);
    LAPPER_CUDA_CHECK(cudaGetLastError());
    uint32_t x = 0;
    LAPPER_CUDA_CHECK(cudaMemcpyAsync(&x, d_out.p, 4, cudaMemcpyDeviceToHost, stream));
    LAPPER_CUDA_CHECK(cudaStreamSynchronize(stream));
    return x;
}

}  // namespace lap
