/*
 * lapper_b200.h -- C ABI of liblapper_b200.so: a B200 (sm_100a) batched interval-overlap engine
 * that is a drop-in for the query path of the Rust crate rust-lapper 1.3.0.
 *
 * The reference has no FFI layer; its boundary is the public Rust API of `Lapper<I,T>`
 * (/root/reference/src/lib.rs:182-680).  Each entry point below names the reference item it
 * replaces.  A Rust shim (`extern "C"` block), the C++ facade
 * (include/lapper_b200.hpp) and the Python/ctypes mirror (rust_lapper_b200/lapper.py) bind
 * exactly these symbols.
 *
 * Conventions
 *  - Coordinates are I = u32 (the reference bench's `Interval<u32, bool>`,
 *    benches/lapper_benchmark.rs:13).  Intervals and queries are half-open [start, stop).
 *  - The payload `val: T` never crosses the ABI: the library returns POSITIONS in the sorted
 *    interval vector (`lapper.intervals`, src/lib.rs:112) plus `perm` (sorted position -> input
 *    position, stable on ties like Vec::sort, src/lib.rs:201), so the caller indexes its own vals.
 *  - Every function returns an int status (LAPPER_OK == 0).  Nothing throws or aborts across the
 *    ABI; lapper_last_error() gives a thread-local message for the last failing call.
 *  - `_dev` variants take DEVICE pointers and a cudaStream_t (passed as void*), enqueue on that
 *    stream and do not synchronise unless stated.  The other variants take HOST pointers (pinned
 *    host memory gives full PCIe speed; pageable works) and return when outputs are complete.
 *  - Inputs are borrowed for the duration of the call; outputs are caller-allocated.
 *  - Threading: concurrent calls that take `const lapper_t*` may run from several host threads on
 *    one handle (the reference's shared `&Lapper`, src/lib.rs:7, :643-645); functions taking a
 *    non-const handle need exclusive access (the reference's `&mut self`).
 *  - There is no CPU fallback: without a CUDA device every build call fails with
 *    LAPPER_ERR_NO_DEVICE.
 */
#ifndef LAPPER_B200_H
#define LAPPER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAPPER_OK 0
#define LAPPER_ERR_INVALID_ARG 1
#define LAPPER_ERR_CUDA 2
#define LAPPER_ERR_OOM 3
#define LAPPER_ERR_PRECONDITION 4 /* input violates start <= stop where the closed form needs it */
#define LAPPER_ERR_NO_DEVICE 5
#define LAPPER_ERR_CAPACITY 6 /* a caller-allocated output buffer is too small; the size needed has been reported */

/* lapper_build_u32_ex flags */
#define LAPPER_BUILD_DEFAULT 0u
#define LAPPER_BUILD_NO_BUCKETS 1u    /* search the sorted arrays directly (sampled keys in smem) */
#define LAPPER_BUILD_FORCE_BUCKETS 2u /* build the bucket-rank table even for tiny sets */
#define LAPPER_BUILD_NO_DENSE 4u      /* no packed (12-bit lane) table in front of the sector table */
#define LAPPER_BUILD_FORCE_DENSE 8u   /* build the packed table whenever the sector table exists */
#define LAPPER_BUILD_DENSE_WIDTH(w) (((uint32_t)(w)) << 16) /* force the packed table with bucket width w, 2 <= w <= 1820 */
#define LAPPER_BUILD_SHIFT(k) ((((uint32_t)(k)) + 1u) << 8) /* force bucket width 2^k, 0 <= k <= 13 */

typedef struct lapper lapper_t;               /* Lapper<u32, T>              src/lib.rs:102-123 */
typedef struct lapper_result lapper_result_t; /* a batch of IterFind outputs src/lib.rs:682-718 */
typedef struct lapper_group lapper_group_t;   /* one replica per GPU, queries sharded            */

int lapper_version(void);
const char *lapper_last_error(void);
int lapper_device_count(int *count_out);
/* number of kernels the count / find entry points have launched in this process so far (measurement aid) */
uint64_t lapper_query_kernel_launches(void);
/* builds with -DLAPPER_DEBUG_CHECKS carry bounds checks in the query kernels (compute-sanitizer is not available on
 * every GPU pool): the number of violated checks on the current device so far; -1 in normal builds */
long long lapper_debug_violations(void);
/* test hook: batches of at least this many queries are offered to the sorted-run kernel (count_sorted.cu), which the
 * batch probe then selects on the device; default 4 Mi (below that the probe's 32 samples say little).  Returns the
 * previous value.  The answers never depend on it. */
uint64_t lapper_debug_set_sorted_min_queries(uint64_t min_queries);
/* runs of 256 queries the sorted-run kernel has answered on the current device since the last reset: out2[0] by
 * ranking the run against the sorted arrays, out2[1] query by query (runs that do not qualify); synchronises the device */
int lapper_debug_sorted_run_stats(uint64_t *out2, int reset);

/* ---- Lapper::new (src/lib.rs:196-236): stable sort by (start, stop), sorted starts, sorted stops,
 *      max_len.  Built on `device` by a radix sort + reductions.  n may be 0. */
int lapper_build_u32(const uint32_t *starts, const uint32_t *stops, uint64_t n, int device, lapper_t **out);
int lapper_build_u32_ex(const uint32_t *starts, const uint32_t *stops, uint64_t n, int device, uint32_t flags,
                        lapper_t **out);
/* inputs already on `device`; enqueued on `stream`, synchronises it before returning */
int lapper_build_u32_dev(const uint32_t *d_starts, const uint32_t *d_stops, uint64_t n, int device, uint32_t flags,
                         void *stream, lapper_t **out);
void lapper_free(lapper_t *l);

/* ---- accessors: Lapper::len/is_empty (src/lib.rs:284-299) and the fields of src/lib.rs:111-122 */
uint64_t lapper_len(const lapper_t *l);
uint32_t lapper_max_len(const lapper_t *l);
int lapper_overlaps_merged(const lapper_t *l);
int lapper_has_inverted(const lapper_t *l); /* 1 if some interval has stop < start */
int lapper_device(const lapper_t *l);
uint32_t lapper_bucket_shift(const lapper_t *l);  /* 0xFFFFFFFF when no bucket table is in use */
uint64_t lapper_index_bytes(const lapper_t *l);   /* device bytes held by the handle */
/* the packed (12-bit lane) table in front of the sector table: bucket width (0 = not built), number of
 * sectors and how many of them overflowed (their queries fall back to the sector table); pointers may be NULL */
int lapper_packed_table_info(const lapper_t *l, uint32_t *width_out, uint64_t *sectors_out, uint64_t *overflowed_out);
/* `intervals` in sorted order as SoA + perm; any pointer may be NULL.  After merge_overlaps perm is
 * the input position of each run's first interval (whose val the reference keeps, src/lib.rs:385-389). */
int lapper_get_intervals(const lapper_t *l, uint32_t *starts_out, uint32_t *stops_out, uint32_t *perm_out);
int lapper_get_starts(const lapper_t *l, uint32_t *out); /* private `starts`, src/lib.rs:114 */
int lapper_get_stops(const lapper_t *l, uint32_t *out);  /* private `stops` (sorted), src/lib.rs:116 */

/* Optional tuning: announce the typical length (stop - start) of the queries to come.  The packed count table --
 * the structure that answers random-order `count` from one 32-byte sector -- is rebuilt with a margin that covers such
 * queries (default: 150-bp reads; useful range up to 640).  Results never change, only which path answers: at the cfg-2
 * shape 100 M random-order 250-bp queries take 0.74 ms instead of 0.94, 400-bp ones 0.84 instead of 1.35.  Not a const
 * call: like merge_overlaps / insert it must not run concurrently with queries on the same handle.  The setting
 * survives merge_overlaps / insert and is copied by lapper_replicate. */
int lapper_tune_query_length(lapper_t *l, uint32_t typical_query_len);

/* Borrowed device pointers, valid until lapper_free / lapper_merge_overlaps. */
typedef struct {
    const uint32_t *starts;       /* sorted starts == intervals[i].start */
    const uint32_t *stops_by_iv;  /* intervals[i].stop */
    const uint32_t *stops_sorted; /* sorted stops */
    const uint32_t *perm;         /* sorted position -> input position */
    const uint32_t *prefix_max;   /* max(stops_by_iv[0..=i]) */
    uint64_t n;
} lapper_dev_view;
int lapper_get_dev_view(const lapper_t *l, lapper_dev_view *out);

/* ---- Lapper::count (src/lib.rs:610-618), batched.  counts_out[j] == lapper.count(q_start[j],
 *      q_stop[j]) as the release-mode `usize` (it wraps exactly where the reference wraps). */
int lapper_count_batch_u32(const lapper_t *l, const uint32_t *q_start, const uint32_t *q_stop, uint64_t nq,
                           uint64_t *counts_out);
/* low 32 bits of the same value (exact whenever the reference does not wrap) */
int lapper_count_batch_u32_c32(const lapper_t *l, const uint32_t *q_start, const uint32_t *q_stop, uint64_t nq,
                               uint32_t *counts_out);
int lapper_count_batch_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop, uint64_t nq,
                               uint32_t *d_counts_out, void *stream);
int lapper_count_batch_u32_dev64(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop, uint64_t nq,
                                 uint64_t *d_counts_out, void *stream);

/* ---- Lapper::find + IterFind::next (src/lib.rs:628-639, :695-718), batched, CSR output:
 *      indices[offsets[j] .. offsets[j+1]) are the positions (in `intervals`) that
 *      lapper.find(q_start[j], q_stop[j]) yields, in iterator order.  Lapper::seek
 *      (src/lib.rs:656-679) yields the same sequence for start-sorted queries; no cursor crosses
 *      the ABI (lapper_seek_cursor below gives the reference's cursor value where a shim wants it). */
int lapper_find_batch_u32(const lapper_t *l, const uint32_t *q_start, const uint32_t *q_stop, uint64_t nq,
                          lapper_result_t **out);
/* same, queries already on the device (enqueued on `stream`, synchronised before returning) */
int lapper_find_batch_u32_devq(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop, uint64_t nq,
                               void *stream, lapper_result_t **out);
uint64_t lapper_result_nq(const lapper_result_t *r);
uint64_t lapper_result_total(const lapper_result_t *r);
/* host copies; either pointer may be NULL.  offsets_out has nq+1 entries, indices_out total entries */
int lapper_result_copy(const lapper_result_t *r, uint64_t *offsets_out, uint32_t *indices_out);
int lapper_result_dev(const lapper_result_t *r, const uint64_t **d_offsets, const uint32_t **d_indices);
void lapper_result_free(lapper_result_t *r);

/* host queries -> host CSR in ONE call, pipelined in chunks (queries in, count + emit, offsets and positions out, all
 * overlapped; pinned buffers give full PCIe speed in both directions at once).  offsets_out has nq + 1 entries and is
 * always written in full, *total_out = offsets_out[nq].  indices_out has room for indices_capacity positions:
 *   - indices_out == NULL: size query (offsets and total only), returns LAPPER_OK;
 *   - total <= indices_capacity: positions written, LAPPER_OK;
 *   - otherwise LAPPER_ERR_CAPACITY with *total_out set (call again with a larger buffer; the positions that were
 *     written are a prefix of the result). */
int lapper_find_batch_u32_host(const lapper_t *l, const uint32_t *q_start, const uint32_t *q_stop, uint64_t nq,
                               uint64_t *offsets_out, uint32_t *indices_out, uint64_t indices_capacity,
                               uint64_t *total_out);

/* ONE call, caller-allocated device buffers of known size (a steady-state pipeline re-uses them): count pass, scan and
 * emit are enqueued back to back -- the emit kernel derives the offsets from the counts itself, so there is no separate
 * offsets pass and no host round trip between the phases; `stream` is synchronised once, at the end, to return the total.
 * d_offsets[nq + 1] is always written in full and *total_out = d_offsets[nq]:
 *   - total <= indices_capacity: d_indices[0 .. total) written, LAPPER_OK;
 *   - otherwise LAPPER_ERR_CAPACITY and d_indices is untouched (call again with a larger buffer, or pass the offsets
 *     to lapper_find_batch_fill_total_u32_dev).
 * d_indices == NULL (indices_capacity 0) is the size query: offsets and total only, LAPPER_OK.  Same results as the
 * two-phase calls below. */
int lapper_find_batch_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop, uint64_t nq,
                              uint64_t *d_offsets, uint32_t *d_indices, uint64_t indices_capacity, uint64_t *total_out,
                              void *stream);

/* two-phase, caller-allocated device buffers: phase 1 writes d_offsets[nq+1] and returns the total
 * (synchronises `stream`); phase 2 fills d_indices[total].  No alignment is required of any buffer (16-byte aligned
 * d_offsets and 32-byte aligned query arrays -- what cudaMalloc and every tensor allocator give -- take the vectorised
 * kernels; anything else takes scalar ones with the same results). */
int lapper_find_batch_offsets_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                      uint64_t nq, uint64_t *d_offsets, uint64_t *total_out, void *stream);
int lapper_find_batch_fill_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                   uint64_t nq, const uint64_t *d_offsets, uint32_t *d_indices, void *stream);
/* same, for a caller that passes phase 1's total back (== d_offsets[nq]): nothing is read back from the device, so
 * the call only enqueues and does not synchronise */
int lapper_find_batch_fill_total_u32_dev(const lapper_t *l, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                         uint64_t nq, const uint64_t *d_offsets, uint64_t total, uint32_t *d_indices,
                                         void *stream);

/* Lapper::seek's cursor update (src/lib.rs:658-671) for one query, so a host shim can keep the
 * reference's `&mut usize` protocol: *cursor is updated exactly as the reference updates it. */
int lapper_seek_cursor(const lapper_t *l, uint32_t start, uint64_t *cursor);

/* ---- Lapper::insert (src/lib.rs:260-273): O(n) like the reference ("very inefficient", :238-239): the sorted
 *      arrays are rebuilt on the device with the new interval at the reference's insert positions
 *      (bsearch_seq_ref: BEFORE any equal (start, stop)), max_len is updated, cov and overlaps_merged are
 *      cleared.  perm of the new interval is *input_position_out = the number of intervals ever given to
 *      this handle before it (so a caller appends its val to its own list). */
int lapper_insert_u32(lapper_t *l, uint32_t start, uint32_t stop, uint64_t *input_position_out);

/* ---- Lapper::merge_overlaps (src/lib.rs:361-416): prefix-max scan, run heads, compaction; rebuilds
 *      starts / stops / max_len; sets overlaps_merged iff the lapper is non-empty.  The query tables of the merged set
 *      are built on its first count / find.
 *      Sets in which some interval has stop < start: the reference's loop (comparisons and checked_sub only) is well
 *      defined for them but is not a prefix-max scan; the engine then runs the loop literally on one warp (about 10 ns
 *      per interval; csrc/merge_sequential.cuh) -- same result as the reference.  (cov, union_and_intersect and depth
 *      refuse such sets with LAPPER_ERR_PRECONDITION because the reference itself overflows on them.) */
int lapper_merge_overlaps(lapper_t *l);
/* ---- Lapper::cov / set_cov (src/lib.rs:310-350) */
int lapper_cov(const lapper_t *l, uint32_t *cov_out);
int lapper_set_cov(lapper_t *l, uint32_t *cov_out);
/* the cached fields `cov: Option<I>` and `overlaps_merged` (src/lib.rs:119-122) as they are, and their restoration
 * when a serialised Lapper (the reference's `with_serde` feature, src/lib.rs:85-86, :1387-1409) is loaded back;
 * any output pointer may be NULL */
int lapper_get_cached_state(const lapper_t *l, int *cov_is_some_out, uint32_t *cov_out, int *overlaps_merged_out);
int lapper_restore_cached_state(lapper_t *l, int cov_is_some, uint32_t cov, int overlaps_merged);
/* ---- Lapper::union_and_intersect / intersect / union (src/lib.rs:512-559) */
int lapper_union_and_intersect(const lapper_t *a, const lapper_t *b, uint32_t *union_out, uint32_t *intersect_out);

/* ---- Lapper::depth() + IterDepth (src/lib.rs:576-598, :720-784): maximal runs of equal coverage depth inside
 *      the merged intervals, in order: (starts[r], stops[r], depth[r]) for r < *n_out.  The three arrays are
 *      malloc'ed by the library (lapper_host_free).  The reference panics on an empty lapper
 *      (src/lib.rs:744): LAPPER_ERR_INVALID_ARG here.  LAPPER_ERR_PRECONDITION if some stop < start. */
int lapper_depth(const lapper_t *l, uint64_t *n_out, uint32_t **starts_out, uint32_t **stops_out, uint32_t **depth_out);
/* frees buffers the library malloc'ed for the caller (lapper_depth, lapper_group_find_batch_u32) */
void lapper_host_free(void *p);

/* ---- multi-GPU, single process: the index is replicated (peer copies over NVLink), queries are
 *      sharded contiguously: device g owns queries [g*nq/G, (g+1)*nq/G).  No data-path collective
 *      for count; find rebases per-shard CSR offsets by the shard totals. */
int lapper_replicate(const lapper_t *l, const int *devices, int g, lapper_group_t **out);
void lapper_group_free(lapper_group_t *grp);
int lapper_group_size(const lapper_group_t *grp);
const lapper_t *lapper_group_replica(const lapper_group_t *grp, int i);
int lapper_group_count_batch_u32(const lapper_group_t *grp, const uint32_t *q_start, const uint32_t *q_stop,
                                 uint64_t nq, uint64_t *counts_out);
int lapper_group_find_batch_u32(const lapper_group_t *grp, const uint32_t *q_start, const uint32_t *q_stop,
                                uint64_t nq, uint64_t *offsets_out /* nq+1 */, uint32_t **indices_out /* lapper_host_free */,
                                uint64_t *total_out);
/* set operations over a group: cov as the sum of the replicas' chunk sums (each replica covers a contiguous chunk of
 * the sorted vector; the cached value of src/lib.rs:311-316 is honoured and set on every replica by _set_cov),
 * merge_overlaps on every replica at once (the replicas stay identical) */
int lapper_group_cov(const lapper_group_t *grp, uint32_t *cov_out);
int lapper_group_set_cov(lapper_group_t *grp, uint32_t *cov_out);
int lapper_group_merge_overlaps(lapper_group_t *grp);

/* ---- index import/export for one-process-per-GPU deployments (torch.distributed broadcast of the
 *      built arrays): build a handle from already-sorted device arrays without re-sorting. */
int lapper_from_sorted_dev(const uint32_t *d_starts, const uint32_t *d_stops_by_iv, const uint32_t *d_stops_sorted,
                           const uint32_t *d_perm, uint64_t n, int overlaps_merged, int device, uint32_t flags,
                           void *stream, lapper_t **out);

/* pinned host memory helpers (cudaHostAlloc / cudaFreeHost) for callers that want full-speed copies */
int lapper_host_alloc_pinned(uint64_t bytes, void **out);
void lapper_host_free_pinned(void *p);

/* =====================================================================================================
 * Lapper<I, T> with I = u64 (src/lib.rs:88-100: the coordinate is any PrimInt + Unsigned; the README's examples use
 * usize).  The second instantiation of the query path (csrc/engine64.cu): the same contracts as the u32 entry points
 * above with 64-bit coordinates -- count wraps in u64 like usize, positions stay u32 (at most 2^32 - 2 intervals),
 * find results come back as the same lapper_result_t.  Plain bisection under a shared-memory key directory instead of
 * the u32 engine's sector tables.  Not instantiated for u64: replica groups, the bincode cache accessors.
 * ===================================================================================================== */
typedef struct lapper64 lapper64_t;

/* Lapper::new (src/lib.rs:196-236); host arrays / device arrays + stream (synchronised before returning) */
int lapper64_build(const uint64_t *starts, const uint64_t *stops, uint64_t n, int device, lapper64_t **out);
int lapper64_build_dev(const uint64_t *d_starts, const uint64_t *d_stops, uint64_t n, int device, void *stream,
                       lapper64_t **out);
void lapper64_free(lapper64_t *l);
uint64_t lapper64_len(const lapper64_t *l);          /* src/lib.rs:284-287 */
uint64_t lapper64_max_len(const lapper64_t *l);      /* src/lib.rs:218-227 */
int lapper64_overlaps_merged(const lapper64_t *l);   /* src/lib.rs:122 */
int lapper64_has_inverted(const lapper64_t *l);
/* the sorted `intervals` vector as SoA + perm (sorted position -> input position); any pointer may be NULL */
int lapper64_get_intervals(const lapper64_t *l, uint64_t *starts_out, uint64_t *stops_out, uint32_t *perm_out);
/* the private sorted `stops` (src/lib.rs:116) */
int lapper64_get_sorted_stops(const lapper64_t *l, uint64_t *stops_sorted_out);

/* Lapper::count (src/lib.rs:610-618), batched; u64 counts wrap like the reference's usize arithmetic */
int lapper64_count_batch(const lapper64_t *l, const uint64_t *q_start, const uint64_t *q_stop, uint64_t nq,
                         uint64_t *counts_out);
int lapper64_count_batch_dev(const lapper64_t *l, const uint64_t *d_q_start, const uint64_t *d_q_stop, uint64_t nq,
                             uint64_t *d_counts_out, void *stream);

/* Lapper::find + IterFind::next (src/lib.rs:628-639, :695-718), batched CSR (see lapper_find_batch_u32 /
 * lapper_find_batch_u32_dev for the contracts, incl. LAPPER_ERR_CAPACITY and the size query) */
int lapper64_find_batch(const lapper64_t *l, const uint64_t *q_start, const uint64_t *q_stop, uint64_t nq,
                        lapper_result_t **out);
int lapper64_find_batch_dev(const lapper64_t *l, const uint64_t *d_q_start, const uint64_t *d_q_stop, uint64_t nq,
                            uint64_t *d_offsets, uint32_t *d_indices, uint64_t indices_capacity, uint64_t *total_out,
                            void *stream);

/* cov / set_cov (src/lib.rs:310-350), merge_overlaps (:361-416), union_and_intersect (:512-559) */
int lapper64_cov(const lapper64_t *l, uint64_t *cov_out);
int lapper64_set_cov(lapper64_t *l, uint64_t *cov_out);
/* `cov: Option<I>` and `overlaps_merged` as they are / restored (the contracts of lapper_get_cached_state and
 * lapper_restore_cached_state): what `with_serde` (src/lib.rs:85-86) needs for a Lapper<u64 | usize, T> */
int lapper64_get_cached_state(const lapper64_t *l, int *cov_is_some_out, uint64_t *cov_out, int *overlaps_merged_out);
int lapper64_restore_cached_state(lapper64_t *l, int cov_is_some, uint64_t cov, int overlaps_merged);
int lapper64_merge_overlaps(lapper64_t *l);
int lapper64_union_and_intersect(const lapper64_t *a, const lapper64_t *b, uint64_t *union_out, uint64_t *intersect_out);
/* insert (src/lib.rs:260-273), the seek cursor (:658-671) and depth (:576-598, :720-784; arrays freed with
 * lapper_host_free, depth values are u32) -- the contracts of lapper_insert_u32 / lapper_seek_cursor / lapper_depth */
int lapper64_insert(lapper64_t *l, uint64_t start, uint64_t stop, uint64_t *input_position_out);
int lapper64_seek_cursor(const lapper64_t *l, uint64_t start, uint64_t *cursor);
int lapper64_depth(const lapper64_t *l, uint64_t *n_out, uint64_t **starts_out, uint64_t **stops_out, uint32_t **depth_out);
/* Narrowing: a u64 set whose coordinates all fit 32 bits (<= 2^32 - 3) has its count / find batches answered by the
 * u32 engine's tables and kernels (same results bit for bit: csrc/engine64.cu, "narrowing"); the u32 index is built on
 * the first batch of >= 4096 queries and dropped by insert / merge_overlaps.  lapper64_is_narrow: 1 once it exists.
 * Sets beyond 32 bits whose span fits (max - min <= 2^32 - 4, e.g. timestamps) are answered the same way on coordinates
 * relative to their smallest one ("rebased form").
 * lapper64_debug_set_narrow_mode (tests): 0 = never, 1 = default, 2 = from the first batch on; returns the old mode. */
int lapper64_is_narrow(const lapper64_t *l);
/* lapper_tune_query_length for a u64 handle (takes effect where the u32 engine answers) */
int lapper64_tune_query_length(lapper64_t *l, uint32_t typical_query_len);
int lapper64_debug_set_narrow_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif
