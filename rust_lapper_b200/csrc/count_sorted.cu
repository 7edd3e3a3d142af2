// count_sorted.cu -- Lapper::count (src/lib.rs:610-618) and the count pass of find_batch for batches whose
// queries come sorted by start AND by stop and lie densely over the intervals (cfg 3: read-sized queries sorted by
// position; the `seek` use case of src/lib.rs:656-679).
//
// The per-query kernels (query.cu) spend ~80 instructions on every query and are issue-bound on such batches.  Here a
// warp takes a CHUNK of consecutive runs of 256 queries and ranks each run against the sorted arrays instead:
//   * the run's starts and stops arrive in shared memory by one bulk copy each (cp.async.bulk + mbarrier, the TMA
//     path: no registers, no LSU instructions; the next run of the warp is in flight while this one is ranked);
//   * the ranks below the run -- lower_bound(starts, x) and upper_bound(stops, y) for an x <= the run's first stop
//     and a y <= its first start -- are CARRIED from the warp's previous run (they are the ranks of that run's last
//     query); only the first run of a chunk, and a run after one that did not qualify, looks them up in the BITS
//     sector table (one sector each, the same address in every lane);
//   * the few dozen starts in [x, b_last) and stops in (y, a_last] are loaded straight from the sorted arrays, one per
//     lane and 64 per side at once (consecutive runs read consecutive keys: nearly always L1 hits); each finds its
//     place among the run's 256 sorted stops (starts) by five shuffle steps over the lanes' last words and three
//     steps in shared memory, and marks it in a 256-entry histogram: +1 for a start, -1 for a stop;
//   * count(a_j, b_j) = lb(x) - ub(y) + sum of the marks at or before j: ONE prefix sum over the histogram gives
//     every query of the run its answer (wrapping u32 like the reference's usize arithmetic; sign-extended for the
//     u64 output, which is exact because the kernel is only offered to sets of fewer than 2^31 intervals).
// A run that is not monotone, ends at u32::MAX, or spans more than 128 starts / stops is answered per query by the
// sector arithmetic of query_views.cuh, so the kernel is exact for any input; it is only LAUNCHED to work when the
// batch probe (query_views.cuh) says the batch has this shape.
// find mode (the count pass of find_batch, src/lib.rs:628-639: BITS count == number of hits when every interval has
// start <= stop and the query is non-empty): the same u32 counts, and the run's total added to its tile's sum.
#include <algorithm>
#include <cstdlib>

#include "query_views.cuh"

namespace lap {
namespace {

#ifndef LAPPER_SORTED_STAGES
#define LAPPER_SORTED_STAGES 2
#endif
#ifndef LAPPER_SORTED_CTAS
#define LAPPER_SORTED_CTAS 4
#endif
#ifndef LAPPER_SORTED_CHUNK
#define LAPPER_SORTED_CHUNK 8
#endif
constexpr int kSThreads = 256;
constexpr int kSWarps = kSThreads / 32;
constexpr int kSRun = (int)kProbeRun;  // queries per run: lane l owns the eight queries [8 l, 8 l + 8)
constexpr int kSStages = LAPPER_SORTED_STAGES;
constexpr uint32_t kSChunk = LAPPER_SORTED_CHUNK;  // consecutive runs a warp takes at a time (ranks carried inside)
constexpr uint32_t kSMaxRounds = 4;    // rounds of 32 starts (stops) per run
static_assert(kSRun == 256, "the 5 + 3 step searches assume runs of 256 queries, eight per lane");

struct __align__(16) WarpArea {
    uint32_t a[kSStages][kSRun];
    uint32_t b[kSStages][kSRun];
    uint32_t hist[kSRun];
    uint64_t full[kSStages];  // mbarriers: the stage's two bulk copies have landed
#if (LAPPER_SORTED_STAGES % 2) == 1
    uint64_t pad_;
#endif
};
constexpr size_t kSSmemBytes = sizeof(WarpArea) * kSWarps;

// runs answered by ranking / per query since the last reset (each warp adds its own tally once, when it is done);
// lapper_debug_sorted_run_stats(): how the tests know which path produced the answers they compare
__device__ unsigned long long g_sorted_stats[2];

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

// #{j : arr[j] <= key} over the run's 256 sorted words; exact when the answer is <= 255.  Lane l holds arr[8 l + 7] in
// `last`: five steps over the lanes by shuffle (no shared-memory traffic, no bank conflicts between the lanes' probes),
// then three steps inside the block of 8 in shared memory.
__device__ __forceinline__ uint32_t run_count_le(const uint32_t *arr, uint32_t last, uint32_t key) {
    uint32_t pos = 0;
#pragma unroll
    for (uint32_t step = 16; step >= 1; step >>= 1) pos += __shfl_sync(0xFFFFFFFFu, last, pos + step - 1) <= key ? step : 0u;
    pos *= 8;
#pragma unroll
    for (uint32_t step = 4; step >= 1; step >>= 1) pos += arr[pos + step - 1] <= key ? step : 0u;
    return pos;
}
// #{j : arr[j] < key}
__device__ __forceinline__ uint32_t run_count_lt(const uint32_t *arr, uint32_t last, uint32_t key) {
    uint32_t pos = 0;
#pragma unroll
    for (uint32_t step = 16; step >= 1; step >>= 1) pos += __shfl_sync(0xFFFFFFFFu, last, pos + step - 1) < key ? step : 0u;
    pos *= 8;
#pragma unroll
    for (uint32_t step = 4; step >= 1; step >>= 1) pos += arr[pos + step - 1] < key ? step : 0u;
    return pos;
}
// where the mark of run position j lives: lane l reads the marks of its queries 8 l .. 8 l + 7 as two conflict-free
// 128-bit words, hist[4 l ..] and hist[128 + 4 l ..]
__device__ __forceinline__ uint32_t hist_slot(uint32_t j) { return ((j & 4u) << 5) + ((j >> 3) << 2) + (j & 3u); }

__device__ __forceinline__ uint32_t warp_inclusive_add(uint32_t x, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    return x;
}

struct SortedArgs {
    BucketView v;
    FindView f;
    const uint32_t *qs, *qe;
    uint64_t nruns;        // runs of kSRun queries handled here: queries [0, nruns * kSRun)
    uint32_t *o32;         // u32 counts (count mode with u32 output; find mode)
    uint64_t *o64;         // count mode, usize output
    uint64_t *tile_sums;   // find mode: sum of the counts per tile of kProbeTile queries (accumulated; zeroed by the caller)
    BatchProbe probe;
};

template <bool kOut64, bool kFind>
__global__ void __launch_bounds__(kSThreads, LAPPER_SORTED_CTAS) count_sorted_kernel(const SortedArgs g) {
    if (!batch_accepted(g.probe)) return;
    extern __shared__ __align__(16) unsigned char ssm_raw[];
    WarpArea &sm = reinterpret_cast<WarpArea *>(ssm_raw)[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp_global = (uint64_t)blockIdx.x * kSWarps + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * kSWarps;
    const BucketView &v = g.v;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kSStages; s++) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    *reinterpret_cast<uint4 *>(sm.hist + 4 * lane) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4 *>(sm.hist + 128 + 4 * lane) = make_uint4(0, 0, 0, 0);
    __syncwarp();
    const uint64_t pol = l2_policy_evict_first();
    // the warp's i-th run: chunk i / kSChunk of its chunks (chunk ids warp_global, warp_global + nwarps, ...), run i % kSChunk
    // of that chunk.  Run ids ascend with i, so the first one out of range ends the warp's work.
    const auto run_at = [&](uint32_t i) -> uint64_t {
        return (warp_global + (uint64_t)(i / kSChunk) * nwarps) * kSChunk + (i % kSChunk);
    };
    const auto fetch = [&](int s, uint64_t run) {
        if (run < g.nruns && lane == 0) {
            mbar_expect_tx(&sm.full[s], 2 * kSRun * 4);
            bulk_load(sm.a[s], g.qs + run * kSRun, kSRun * 4, &sm.full[s], pol);
            bulk_load(sm.b[s], g.qe + run * kSRun, kSRun * 4, &sm.full[s], pol);
        }
    };
#pragma unroll
    for (int s = 0; s < kSStages; s++) fetch(s, run_at((uint32_t)s));
    uint32_t tally_fast = 0, tally_slow = 0;
    // ranks carried from the previous run: r_next = #{starts < x_s}, u_next = #{stops <= x_e}
    bool have = false;
    uint32_t r_next = 0, u_next = 0, x_s = 0, x_e = 0;
    for (uint32_t i0 = 0;; i0 += kSStages) {
#pragma unroll
        for (int s = 0; s < kSStages; s++) {
            const uint32_t i = i0 + (uint32_t)s;
            const uint64_t run = run_at(i);
            if (run >= g.nruns) {
                if (lane == 0 && (tally_fast | tally_slow)) {
                    atomicAdd(&g_sorted_stats[0], (unsigned long long)tally_fast);
                    atomicAdd(&g_sorted_stats[1], (unsigned long long)tally_slow);
                }
                return;
            }
            mbar_wait(&sm.full[s], (i0 / kSStages) & 1u);
            const uint32_t *A = sm.a[s], *B = sm.b[s];
            // the lane owns queries 8 lane .. 8 lane + 7 of the run
            uint32_t a7, b7;
            bool ok;
            uint32_t a_first, b_first;
            {
                const uint4 a0 = *reinterpret_cast<const uint4 *>(A + 8 * lane), a1 = *reinterpret_cast<const uint4 *>(A + 8 * lane + 4);
                const uint4 b0 = *reinterpret_cast<const uint4 *>(B + 8 * lane), b1 = *reinterpret_cast<const uint4 *>(B + 8 * lane + 4);
                // both coordinates non-decreasing over the run?
                uint32_t a_nx = __shfl_down_sync(0xFFFFFFFFu, a0.x, 1), b_nx = __shfl_down_sync(0xFFFFFFFFu, b0.x, 1);
                if (lane == 31) a_nx = b_nx = 0xFFFFFFFFu;
                ok = a0.x <= a0.y && a0.y <= a0.z && a0.z <= a0.w && a0.w <= a1.x && a1.x <= a1.y && a1.y <= a1.z &&
                     a1.z <= a1.w && a1.w <= a_nx && b0.x <= b0.y && b0.y <= b0.z && b0.z <= b0.w && b0.w <= b1.x &&
                     b1.x <= b1.y && b1.y <= b1.z && b1.z <= b1.w && b1.w <= b_nx;
                if (kFind) {
                    // BITS count is the number of hits only for a non-empty query
                    ok = ok && a0.x < b0.x && a0.y < b0.y && a0.z < b0.z && a0.w < b0.w && a1.x < b1.x && a1.y < b1.y &&
                         a1.z < b1.z && a1.w < b1.w;
                }
                a7 = a1.w, b7 = b1.w;
                a_first = __shfl_sync(0xFFFFFFFFu, a0.x, 0), b_first = __shfl_sync(0xFFFFFFFFu, b0.x, 0);
            }
            const uint32_t a_last = __shfl_sync(0xFFFFFFFFu, a7, 31), b_last = __shfl_sync(0xFFFFFFFFu, b7, 31);
            ok = ok && a_last != 0xFFFFFFFFu;  // src/lib.rs:614: `start + 1` wraps there
            bool fast = __all_sync(0xFFFFFFFFu, ok);
            uint32_t out[8];
            if (fast) {
                // ranks below the run: carried from the warp's previous run when this one continues it, else one
                // sector each (the same address in every lane)
                uint32_t r_lo, u_lo;
                if (have && (i % kSChunk) != 0 && b_first >= x_s && a_first >= x_e) {
                    r_lo = r_next, u_lo = u_next;
                } else {
                    const Sector sb = ld_sector_cached(v.buckets + bucket_of(v, b_first));
                    const Sector sa = ld_sector_cached(v.buckets + bucket_of(v, a_first));
                    r_lo = lb_from_sector(v, sb, b_first);
                    u_lo = ub_from_sector(v, sa, a_first);
                }
                // 64 keys per side at once: round 0 and, for the one run in seven that needs it, round 1 (for the others
                // it is the L1 prefetch of the next run's round 0)
                const uint32_t rem_s = v.n - r_lo, rem_e = v.n - u_lo;
                const uint32_t ks0 = lane < rem_s ? __ldg(v.S + r_lo + lane) : 0xFFFFFFFFu;
                const uint32_t ke0 = lane < rem_e ? __ldg(v.E + u_lo + lane) : 0xFFFFFFFFu;
                const uint32_t ks1 = lane + 32u < rem_s ? __ldg(v.S + r_lo + 32u + lane) : 0xFFFFFFFFu;
                const uint32_t ke1 = lane + 32u < rem_e ? __ldg(v.E + u_lo + 32u + lane) : 0xFFFFFFFFu;
                // start i is below b_j exactly for the j at or after its place among the b's; stop i is <= a_j exactly for
                // the j at or after its place among the a's.  A padding key (u32::MAX) is never valid: b_last <= MAX, a_last < MAX.
                uint32_t n_s, n_e;
                {
                    const bool vs = ks0 < b_last, ve = ke0 <= a_last;
                    const uint32_t ps = run_count_le(B, b7, ks0), pe = run_count_lt(A, a7, ke0);
                    if (vs) atomicAdd(&sm.hist[hist_slot(ps)], 1u);
                    if (ve) atomicAdd(&sm.hist[hist_slot(pe)], 0xFFFFFFFFu);
                    n_s = __popc(__ballot_sync(0xFFFFFFFFu, vs)), n_e = __popc(__ballot_sync(0xFFFFFFFFu, ve));
                }
                if (n_s == 32u) {
                    for (uint32_t r = 1;; r++) {
                        if (r == kSMaxRounds) {
                            fast = false;
                            break;
                        }
                        const uint32_t k = 32u * r + lane;
                        const uint32_t key = r == 1 ? ks1 : (k < rem_s ? __ldg(v.S + r_lo + k) : 0xFFFFFFFFu);
                        const bool vs = key < b_last;
                        const uint32_t ps = run_count_le(B, b7, key);
                        if (vs) atomicAdd(&sm.hist[hist_slot(ps)], 1u);
                        const uint32_t c = __popc(__ballot_sync(0xFFFFFFFFu, vs));
                        n_s += c;
                        if (c != 32u) break;
                    }
                }
                if (n_e == 32u && fast) {
                    for (uint32_t r = 1;; r++) {
                        if (r == kSMaxRounds) {
                            fast = false;
                            break;
                        }
                        const uint32_t k = 32u * r + lane;
                        const uint32_t key = r == 1 ? ke1 : (k < rem_e ? __ldg(v.E + u_lo + k) : 0xFFFFFFFFu);
                        const bool ve = key <= a_last;
                        const uint32_t pe = run_count_lt(A, a7, key);
                        if (ve) atomicAdd(&sm.hist[hist_slot(pe)], 0xFFFFFFFFu);
                        const uint32_t c = __popc(__ballot_sync(0xFFFFFFFFu, ve));
                        n_e += c;
                        if (c != 32u) break;
                    }
                }
                __syncwarp();
                uint4 h0 = *reinterpret_cast<const uint4 *>(sm.hist + 4 * lane), h1 = *reinterpret_cast<const uint4 *>(sm.hist + 128 + 4 * lane);
                *reinterpret_cast<uint4 *>(sm.hist + 4 * lane) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(sm.hist + 128 + 4 * lane) = make_uint4(0, 0, 0, 0);
                if (fast) {
                    // marks at or before each of the lane's queries
                    h0.y += h0.x, h0.z += h0.y, h0.w += h0.z;
                    h1.x += h0.w, h1.y += h1.x, h1.z += h1.y, h1.w += h1.z;
                    const uint32_t before = warp_inclusive_add(h1.w, lane) - h1.w + (r_lo - u_lo);
                    out[0] = before + h0.x, out[1] = before + h0.y, out[2] = before + h0.z, out[3] = before + h0.w;
                    out[4] = before + h1.x, out[5] = before + h1.y, out[6] = before + h1.z, out[7] = before + h1.w;
                    r_next = r_lo + n_s, u_next = u_lo + n_e, x_s = b_last, x_e = a_last;
                    have = true;
                }
            }
            if (!fast) {
                // a run that does not qualify: every query on its own by the sector arithmetic (src/lib.rs:611-617; in
                // find mode the hits by the rule that is valid for the query).  The loop is not unrolled -- eight copies
                // of that code would only bloat the rare path -- so the answers travel through the lane's own words of
                // the stage instead of dynamically indexed registers.
                uint32_t *const Aw = sm.a[s];
                __syncwarp();
#pragma unroll 1
                for (uint32_t k = 0; k < 8; k++) {
                    const uint32_t j = 8u * lane + k;
                    const uint32_t qa = Aw[j], qb = B[j];
                    uint32_t l, u = 0;
                    if (kFind)
                        l = find_count_one(g.f, qa, qb);
                    else
                        bits_count_slow(v, qa, qb, l, u);
                    Aw[j] = l - u;
                }
                const uint4 l0 = *reinterpret_cast<const uint4 *>(Aw + 8 * lane), l1 = *reinterpret_cast<const uint4 *>(Aw + 8 * lane + 4);
                out[0] = l0.x, out[1] = l0.y, out[2] = l0.z, out[3] = l0.w, out[4] = l1.x, out[5] = l1.y, out[6] = l1.z, out[7] = l1.w;
                have = false;
            }
            tally_fast += fast ? 1u : 0u, tally_slow += fast ? 0u : 1u;
            __syncwarp();  // every lane is done with the stage
            fetch(s, run_at(i + kSStages));  // refill the stage with the warp's next run but kSStages - 1
            const uint64_t q0 = run * kSRun + 8 * lane;  // the lane's first query
            if (kOut64) {
                // len - first - (len - last) in wrapping usize == last - first (mod 2^64); |last - first| <= n < 2^31
                Sector o;
#pragma unroll
                for (int h = 0; h < 2; h++) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        o.w[2 * k] = out[4 * h + k], o.w[2 * k + 1] = (uint32_t)((int32_t)out[4 * h + k] >> 31);
                    }
                    st_stream_v8(g.o64 + q0 + 4 * h, o);
                }
            } else {
                Sector o;
#pragma unroll
                for (int k = 0; k < 8; k++) o.w[k] = out[k];
                st_stream_v8(g.o32 + q0, o);
            }
            if (kFind) {
                // the run's share of its tile's sum
                const uint32_t any = out[0] | out[1] | out[2] | out[3] | out[4] | out[5] | out[6] | out[7];
                unsigned long long total;
                if (!__any_sync(0xFFFFFFFFu, any >> 24)) {  // 256 counts below 2^24: the sum fits 32 bits
                    total = __reduce_add_sync(0xFFFFFFFFu, out[0] + out[1] + out[2] + out[3] + out[4] + out[5] + out[6] + out[7]);
                } else {
                    total = (unsigned long long)out[0] + out[1] + out[2] + out[3] + out[4] + out[5] + out[6] + out[7];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
                }
                if (lane == 0 && total)
                    atomicAdd(reinterpret_cast<unsigned long long *>(g.tile_sums + run / (kProbeTile / kProbeRun)), total);
            }
        }
    }
}

template <typename K>
uint32_t sorted_grid(K kernel) {
    // once per device: room for the runs in flight, and the number of CTAs that are resident together
    static std::atomic<int> per_sm[64];
    int dev = 0;
    LAPPER_CUDA_CHECK(cudaGetDevice(&dev));
    int r = (dev >= 0 && dev < 64) ? per_sm[dev].load(std::memory_order_acquire) : 0;
    if (r == 0) {
        LAPPER_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSSmemBytes));
        LAPPER_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        LAPPER_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, kernel, kSThreads, kSSmemBytes));
        r = std::max(r, 1);
        if (dev >= 0 && dev < 64) per_sm[dev].store(r, std::memory_order_release);
    }
    return sm_count() * (uint32_t)r;
}

}  // namespace

void sorted_run_stats(uint64_t out[2], bool reset) {
    unsigned long long v[2] = {0, 0};
    LAPPER_CUDA_CHECK(cudaDeviceSynchronize());
    LAPPER_CUDA_CHECK(cudaMemcpyFromSymbol(v, g_sorted_stats, sizeof(v)));
    out[0] = v[0], out[1] = v[1];
    if (reset) {
        const unsigned long long z[2] = {0, 0};
        LAPPER_CUDA_CHECK(cudaMemcpyToSymbol(g_sorted_stats, z, sizeof(z)));
    }
}

bool count_sorted_disabled() {
    static const bool off = [] {
        const char *e = std::getenv("LAPPER_NO_SORTED");
        return e && atoi(e) != 0;
    }();
    return off;
}

// queries [0, nruns * 256) (the caller's other kernels handle the rest); exits at once unless the batch probe accepts.
// tile_sums != nullptr selects find mode (u32 counts + per-tile sums).
void launch_count_sorted(const DeviceIndex &ix, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nruns, uint32_t *d_out32,
                         uint64_t *d_out64, uint64_t *tile_sums, const BatchProbe &probe, cudaStream_t stream) {
    SortedArgs g;
    g.v = make_bucket_view(ix);
    g.f = make_find_view(ix);
    g.qs = d_qs, g.qe = d_qe, g.nruns = nruns;
    g.o32 = d_out32, g.o64 = d_out64, g.tile_sums = tile_sums;
    g.probe = probe;
    const uint64_t want = div_up(div_up(nruns, kSChunk), kSWarps);
    const auto launch = [&](auto kernel) {
        const uint32_t grid = (uint32_t)std::min<uint64_t>(want, sorted_grid(kernel));
        g_query_launches++, kernel<<<grid, kSThreads, kSSmemBytes, stream>>>(g);
    };
    if (tile_sums)
        launch(count_sorted_kernel<false, true>);
    else if (d_out64)
        launch(count_sorted_kernel<true, false>);
    else
        launch(count_sorted_kernel<false, false>);
    LAPPER_CUDA_CHECK(cudaGetLastError());
}

}  // namespace lap
