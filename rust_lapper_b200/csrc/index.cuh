// index.cuh -- the device-resident index behind a lapper_t, and the launchers each .cu exports.
//
// HBM layout (all u32, SoA; the reference's AoS Vec<Interval<I,T>> is a CPU artefact):
//   starts[n]        intervals[i].start, i in sorted (start, stop) order      (src/lib.rs:112,114)
//   stops_by_iv[n]   intervals[i].stop                                        (src/lib.rs:112)
//   stops_sorted[n]  the reference's private sorted `stops`                   (src/lib.rs:116)
//   perm[n]          sorted position -> input position (stable on ties)
//   prefix_max[n]    max(stops_by_iv[0..=i]); drives find's window, merge_overlaps and cov
//   buckets[nb+1]    32-byte BITS sectors, see below
#pragma once

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "packed_sector.cuh"

namespace lap {

// ---- BITS bucket sector -------------------------------------------------------------------------
// Bucket B covers coordinates [base, base + w), base = B << shift, w = 2^shift.  One 32-byte sector
// (one LDG.E.256, one L2 sector, one L1 wavefront) answers both halves of BITS count for a query
// whose stop falls in the bucket and whose start is not more than the margin m = w/4 below it:
//   w[0]      rs    = #starts        <  base
//   w[1]      re_hi = #sorted stops  <  base + w
//   w[2..7]   12 x 16-bit lanes, each with guard bit 15 set:
//               a start s in [base, base + w)       ->  0x8000 | (s - base)
//               a stop  e in [base - m, base + w)   ->  0x8000 | (w + (e - base + m))
//               empty lane                          ->  0x8000 | w
//   lower_bound(starts, b) = rs    + 12 - #{lane >= b - base}            for b in the bucket
//   upper_bound(stops,  a) = re_hi      - #{lane >= w + (a - base + m) + 1}  for a in [base - m, base + w)
// Both lane counts are one packed 16-bit subtract per word: (0x8000|v) - x keeps bit 15 iff v >= x and
// never borrows across lanes because x <= 2w + m + 1 < 0x8000 (shift <= 13).
// A bucket whose lanes would hold more than 12 keys is marked (w[7] = 0xFFFFFFFF) and refined: w[2]
// points at 2^w[3] child sectors in `pool`, each a sector of the same format for a 2^(shift - w[3])
// wide slice of the bucket (margin 0); w[4] = #starts < base + w and w[5] = #stops < base bound the
// bucket's slices of the sorted arrays for the rare child that still overflows (many equal
// coordinates).  Sector nb is a sentinel for coordinates beyond the last key: rs = re_hi = n, no keys of its
// own, only the stops of its margin.
constexpr uint32_t kBucketLanes = 12;
constexpr uint32_t kMaxBucketShift = 13;
constexpr uint32_t kLaneGuard = 0x80008000u;
constexpr uint32_t kNoChild = 0xFFFFFFFFu;

// The packed ("dense") table in front of this one: packed_sector.cuh.

// histogram of interval lengths by bit length (bin 0 = length 0 or inverted) + the longest per bin, and the other
// scalars of one pass over the sorted intervals (finish_index_core)
struct BuildStats {
    unsigned long long count[33];
    uint32_t maxlen[33];
    uint32_t max_len;      // max over intervals of stop.checked_sub(start).unwrap_or(0)   (src/lib.rs:218-227)
    uint32_t inverted;     // some interval has stop < start
    uint32_t first_start, last_start, last_stop;  // S[0], S[n-1], sorted stops [n-1]
    uint32_t max_perm;     // max input position
};

struct DeviceIndex {
    uint64_t n = 0;
    uint32_t *starts = nullptr;
    uint32_t *stops_by_iv = nullptr;
    uint32_t *stops_sorted = nullptr;
    uint32_t *perm = nullptr;
    uint32_t *prefix_max = nullptr;
    Sector *buckets = nullptr;
    uint64_t nbuckets = 0;  // real buckets; the array has nbuckets + 1 sectors
    Sector *pool = nullptr;  // child sectors of overflowing buckets
    uint64_t npool = 0;
    uint32_t margin = 0;     // stops this far below a bucket are also in its lanes
    Sector *dense = nullptr;  // packed sectors, dense_nb + 1 of them (the last one is the sentinel)
    uint64_t dense_nb = 0;
    uint32_t dense_w = 0, dense_m = 0;
    uint32_t dense_magic = 0;  // ceil(2^32 / dense_w)
    uint64_t dense_overflow = 0;  // sectors that did not fit (their queries fall back)
    uint32_t query_len_hint = 0;  // typical query length the caller announced (lapper_tune_query_length): the packed table's margin covers it
    uint32_t shift = 0xFFFFFFFFu;
    uint32_t max_len = 0;
    uint32_t max_coord = 0;
    bool has_inverted = false;
    uint64_t bytes = 0;
    // the query tables (buckets / pool / dense, length classes, blkmax) are derived data: built at once by
    // lapper_build_*, lazily -- on the first count / find -- after merge_overlaps and insert (build_query_tables)
    bool tables_ready = false;
    BuildStats stats = {};  // host copy (finish_index_core)

    // ---- find: length classes (build.cu::build_find_classes).  The sorted intervals are split,
    // order-preserving, into up to kMaxClasses groups by length; class c occupies
    // [cls_off[c], cls_off[c] + cls_n[c]) of cls_starts / cls_stops / cls_gidx (gidx = position in the
    // global sorted order).  Inside a class no interval is longer than cls_maxlen[c], so
    // find(a, b) only has to look at starts in [a - cls_maxlen[c], b) of that class -- a window that is
    // short for the short classes and dense in hits for the long ones.  cls_dir[B] holds, for the
    // coordinate bucket B (width 2^cls_shift), the rank of B's first coordinate in each class.
    uint32_t ncls = 0;
    uint32_t cls_off[4] = {0, 0, 0, 0};
    uint32_t cls_n[4] = {0, 0, 0, 0};
    uint32_t cls_maxlen[4] = {0, 0, 0, 0};
    uint32_t *cls_starts = nullptr;
    uint32_t *cls_stops = nullptr;
    uint32_t *cls_gidx = nullptr;
    uint4 *cls_dir = nullptr;  // cls_nb + 1 entries

    // ---- find, heavy queries: max(stops_by_iv) over blocks of 32 and of 1024 consecutive sorted intervals.
    // A warp that scans one query's window skips every block whose max stop is <= the query start (the
    // `max_ends` idea of the reference's devlog.md:2-5).
    uint32_t *blkmax32 = nullptr;    // ceil(n / 32)
    uint32_t *blkmax1024 = nullptr;  // ceil(n / 1024)
    uint32_t cls_shift = 0;
    uint64_t cls_nb = 0;
};
constexpr uint32_t kMaxClasses = 4;

}  // namespace lap

// The opaque handle of the C ABI.
struct lapper {
    int device = 0;
    lap::DeviceIndex ix;
    std::mutex tables_mu;  // serialises the lazy build of ix's query tables between concurrent const calls
    std::atomic<bool> tables_flag{false};  // ix.tables_ready, readable without the mutex
    bool overlaps_merged = false;
    bool cov_is_some = false;
    uint32_t cov = 0;
    uint32_t build_flags = 0;
    uint64_t next_input_id = 0;  // input position handed to the next inserted interval
    // hits per query (x 256, rounded up) of the last find_batch on this handle; 0: none yet.  find_batch sizes its
    // position buffer from it and runs both phases in one enqueue (launch_find)
    mutable std::atomic<uint64_t> find_hits_per_query_x256{0};
};

struct lapper_result {  // device CSR of one find_batch call (pool memory)
    int device = 0;
    uint64_t nq = 0;
    uint64_t total = 0;
    uint64_t *d_offsets = nullptr;
    uint32_t *d_indices = nullptr;
};

namespace lap {

// build.cu ---------------------------------------------------------------------------------------
// Sort (start, stop) pairs (stable, by (start, stop)) and derive every array of DeviceIndex.
void build_index_from_unsorted(DeviceIndex &ix, const uint32_t *d_starts, const uint32_t *d_stops, uint64_t n,
                               uint32_t flags, cudaStream_t stream);
// Arrays already sorted (import / after merge_overlaps / insert); ix owns the four sorted arrays (pool memory).
// finish_index_core: max_len, has_inverted, max_coord, prefix_max (what cov / merge / the accessors need; one host
// sync).  build_query_tables: everything count / find need on top of that (length classes, blkmax, BITS sectors,
// packed table); sets ix.tables_ready.
void finish_index_core(DeviceIndex &ix, cudaStream_t stream);
void build_query_tables(DeviceIndex &ix, uint32_t flags, cudaStream_t stream);
// (re)builds the packed count table alone, for the current ix.query_len_hint (the other tables must exist)
void rebuild_packed_table(DeviceIndex &ix, uint32_t flags, cudaStream_t stream);
// every array of the index goes back to the library pool, stream-ordered on `stream`
void free_index(DeviceIndex &ix, cudaStream_t stream);
// index memory: stream-ordered allocations from the library pool (common.cuh) -- a rebuilt index reuses the
// memory of the one it replaces instead of going through cudaMalloc / cudaFree (milliseconds per call at 50 M)
template <typename T>
T *pool_alloc(uint64_t count, uint64_t &bytes, cudaStream_t stream) {
    T *p = nullptr;
    const uint64_t b = (count ? count : 1) * sizeof(T);
    LAPPER_CUDA_CHECK(cudaMallocFromPoolAsync((void **)&p, b, scratch_pool(), stream));
    bytes += b;
    return p;
}
inline void pool_free(void *p, cudaStream_t stream) {
    if (p) cudaFreeAsync(p, stream);
}
// in-place ascending sort of n u32 keys (hand-written LSD radix sort); tmp holds n keys
void launch_sort_u32(uint32_t *keys, uint32_t *tmp, uint64_t n, cudaStream_t stream);
// the same sort on u64 keys, alone or with u32 payloads (stable: equal keys keep their order); tmp arrays hold n items
void launch_sort_u64(uint64_t *keys, uint64_t *tmp, uint64_t n, cudaStream_t stream);
void launch_sort_pairs_u64(uint64_t *keys, uint64_t *keys_tmp, uint32_t *vals, uint32_t *vals_tmp, uint64_t n, cudaStream_t stream);

// query.cu ---------------------------------------------------------------------------------------
extern std::atomic<uint64_t> g_query_launches;  // kernels launched by launch_count / launch_find_* in this process
extern std::atomic<uint64_t> g_sorted_min_queries;  // lapper_debug_set_sorted_min_queries()
long long debug_violations();
// count_sorted.cu: runs of 256 queries answered by ranking / per query on the current device since the last reset
void sorted_run_stats(uint64_t out[2], bool reset);  // LAPPER_DEBUG_CHECKS builds: violated LAPPER_CHECKs so far; -1 otherwise
void launch_count(const DeviceIndex &ix, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq, uint32_t *d_out32,
                  uint64_t *d_out64, cudaStream_t stream);
// writes d_offsets[nq+1]; returns total hits (synchronises the stream)
uint64_t launch_find_offsets(const DeviceIndex &ix, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq,
                             uint64_t *d_offsets, cudaStream_t stream);
// one enqueue for both phases (query.cu): returns the total; d_indices is written only when the total fits `capacity`
uint64_t launch_find(const DeviceIndex &ix, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq, uint64_t *d_offsets,
                     uint32_t *d_indices, uint64_t capacity, cudaStream_t stream);
// total = offsets[nq] when the caller knows it (saves a device round trip), kTotalUnknown otherwise
constexpr uint64_t kTotalUnknown = ~0ull;
void launch_find_fill(const DeviceIndex &ix, const uint32_t *d_qs, const uint32_t *d_qe, uint64_t nq,
                      const uint64_t *d_offsets, uint32_t *d_indices, cudaStream_t stream,
                      uint64_t total = kTotalUnknown);
// out[j] = in[j] + base for j < count (CSR offsets of a shard / chunk -> offsets of the whole batch)
void launch_rebase_offsets(const uint64_t *in, uint64_t *out, uint64_t count, uint64_t base, cudaStream_t stream);

// setops.cu --------------------------------------------------------------------------------------
// order-preserving split of the sorted intervals into length classes (fills ix.cls_starts/stops/gidx)
void launch_class_partition(DeviceIndex &ix, const uint8_t *class_of_bin /* host, 33 entries */, cudaStream_t stream);
// exclusive prefix sum (u32, wrapping) of `in` into `out` (may alias)
void launch_exclusive_sum(const uint32_t *in, uint32_t *out, uint64_t n, cudaStream_t stream);
// inclusive prefix max of `in` (n elements) into `out`
void launch_prefix_max(const uint32_t *in, uint32_t *out, uint64_t n, cudaStream_t stream);
uint32_t compute_cov(const DeviceIndex &ix, cudaStream_t stream);
// the share of cov contributed by the sorted positions [lo, hi) (wrapping u32; the shares of a partition add up to cov)
uint32_t compute_cov_range(const DeviceIndex &ix, uint64_t lo, uint64_t hi, cudaStream_t stream);
// merged (start, stop, head perm) arrays; returns m.  Outputs are pool memory (pool_free; m may be 0).
uint64_t compute_merged(const DeviceIndex &ix, uint32_t **m_starts, uint32_t **m_stops, uint32_t **m_perm,
                        cudaStream_t stream);
// depth(): maximal runs of equal coverage depth (src/lib.rs:576-598, :735-784).  Device arrays (pool memory,
// pool_free), returns the number of runs.
uint64_t compute_depth(const DeviceIndex &ix, uint32_t **d_starts, uint32_t **d_stops, uint32_t **d_depth,
                       cudaStream_t stream);
// measure of (union of a) intersected with (union of b), wrapping u32
uint32_t compute_intersect_measure(const DeviceIndex &a, const DeviceIndex &b, cudaStream_t stream);

}  // namespace lap
