/*
 * lapper_oracle.h -- CPU restatement of rust-lapper 1.3.0 (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle and the reported CPU baseline.  It is NOT part of
 * the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path
 * (rust_lapper_b200/csrc) never links or calls anything in this directory.
 *
 * Every function restates one function of /root/reference/src/lib.rs with the
 * same data layout (AoS Vec<Interval>, sorted `starts`, sorted `stops`,
 * `max_len`) and the same loop structure; the file:line it follows is cited
 * on each declaration.  Coordinates are I = u32 (the reference bench's type,
 * benches/lapper_benchmark.rs:13) or, built with -DLO_COORD_BITS=64, I = u64; `usize` results are uint64_t and wrap the
 * way a release build of the reference wraps.
 *
 * Parity pin: the reference is Rust and no Rust toolchain exists in this
 * image, so oracle/_ref cannot be built.  The oracle is pinned instead against
 * every known-answer test of src/lib.rs:851-1411, the doctests and
 * examples/ex1.rs:65-121 (tests/test_oracle_golden.py).
 */
#ifndef LAPPER_ORACLE_H
#define LAPPER_ORACLE_H

#include <stddef.h>
#include <stdint.h>

/* I, the coordinate type (src/lib.rs:88-100: PrimInt + Unsigned): u32 by default, u64 with -DLO_COORD_BITS=64
 * (liblapper_oracle64.so, the checker of the u64 engine).  val, positions and depths stay 32-bit. */
#ifndef LO_COORD_BITS
#define LO_COORD_BITS 32
#endif
#if LO_COORD_BITS == 64
typedef uint64_t lo_coord;
#else
typedef uint32_t lo_coord;
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* src/lib.rs:88-100  Interval<I,T>{start,stop,val}; val stands in for T and is
 * ignored by ordering and equality (src/lib.rs:145-180). */
typedef struct {
    lo_coord start;
    lo_coord stop;
    uint32_t val;
} lo_interval;

/* src/lib.rs:102-123  Lapper<I,T> */
typedef struct {
    lo_interval *intervals; /* sorted by (start, stop), stable */
    lo_coord *starts;       /* sorted starts */
    lo_coord *stops;        /* independently sorted stops */
    size_t len;
    size_t cap;
    lo_coord max_len;
    int cov_is_some; /* Option<I> tag */
    lo_coord cov;
    int overlaps_merged;
} lo_lapper;

/* src/lib.rs:682-693  IterFind */
typedef struct {
    const lo_lapper *inner;
    size_t off;
    lo_coord start;
    lo_coord stop;
} lo_iter_find;

/* src/lib.rs:720-733  IterDepth */
typedef struct {
    const lo_lapper *inner;
    lo_lapper *merged;
    lo_coord curr_merged_pos;
    size_t curr_pos;
    size_t cursor;
    size_t end;
} lo_iter_depth;

/* src/lib.rs:196-236  Lapper::new.  vals may be NULL (val = input position). */
lo_lapper *lo_new(const lo_coord *starts, const lo_coord *stops, const uint32_t *vals, size_t n);
void lo_free(lo_lapper *l);

/* src/lib.rs:260-273  Lapper::insert */
void lo_insert(lo_lapper *l, lo_coord start, lo_coord stop, uint32_t val);

/* src/lib.rs:130-136  Interval::intersect */
lo_coord lo_interval_intersect(lo_coord s1, lo_coord e1, lo_coord s2, lo_coord e2);
/* src/lib.rs:138-142  Interval::overlap */
int lo_interval_overlap(const lo_interval *iv, lo_coord start, lo_coord stop);

/* src/lib.rs:423-437  Lapper::lower_bound */
size_t lo_lower_bound(lo_coord start, const lo_interval *intervals, size_t n);
/* src/lib.rs:448-466  Lapper::bsearch_seq_ref on a key slice */
size_t lo_bsearch_seq(lo_coord key, const lo_coord *elems, size_t n);

/* src/lib.rs:611-618  Lapper::count (usize, release-mode wrapping) */
uint64_t lo_count(const lo_lapper *l, lo_coord start, lo_coord stop);

/* src/lib.rs:629-639  Lapper::find */
lo_iter_find lo_find(const lo_lapper *l, lo_coord start, lo_coord stop);
/* src/lib.rs:657-679  Lapper::seek */
lo_iter_find lo_seek(const lo_lapper *l, lo_coord start, lo_coord stop, size_t *cursor);
/* src/lib.rs:704-717  IterFind::next; returns the position in l->intervals, or -1 for None */
int64_t lo_iter_find_next(lo_iter_find *it);

/* src/lib.rs:327-350 / 311-316 / 320-324 */
lo_coord lo_calculate_coverage(const lo_lapper *l);
lo_coord lo_cov(const lo_lapper *l);
lo_coord lo_set_cov(lo_lapper *l);

/* src/lib.rs:363-416 */
void lo_merge_overlaps(lo_lapper *l);

/* src/lib.rs:513-545 */
void lo_union_and_intersect(const lo_lapper *self, const lo_lapper *other, lo_coord *union_out,
                            lo_coord *intersect_out);

/* src/lib.rs:577-598 / 743-783.  lo_depth aborts on an empty lapper like the reference panics. */
lo_iter_depth lo_depth(const lo_lapper *l);
int lo_iter_depth_next(lo_iter_depth *it, lo_coord *start, lo_coord *stop, uint32_t *val);
void lo_iter_depth_free(lo_iter_depth *it);

/* ---- batch drivers (loops of the per-query API, as benches/lapper_benchmark.rs:102-124 loops it) ---- */

/* counts_out[j] = lo_count(q_start[j], q_stop[j]); nthreads > 1 shards queries contiguously over
 * pthreads sharing one immutable lapper (src/lib.rs:7). */
void lo_count_batch(const lo_lapper *l, const lo_coord *q_start, const lo_coord *q_stop, size_t nq,
                    uint64_t *counts_out, int nthreads);

/* Pass 1: offsets_out[nq+1] (exclusive prefix of per-query find().count()).  Returns total hits. */
uint64_t lo_find_batch_offsets(const lo_lapper *l, const lo_coord *q_start, const lo_coord *q_stop,
                               size_t nq, uint64_t *offsets_out, int nthreads);
/* Pass 2: indices_out[offsets[j] .. offsets[j+1]) = positions yielded by find(q_j), iterator order. */
void lo_find_batch_fill(const lo_lapper *l, const lo_coord *q_start, const lo_coord *q_stop, size_t nq,
                        const uint64_t *offsets, uint32_t *indices_out, int nthreads);
/* Same via seek with one cursor per thread shard (valid for start-sorted queries). */
void lo_seek_batch_fill(const lo_lapper *l, const lo_coord *q_start, const lo_coord *q_stop, size_t nq,
                        const uint64_t *offsets, uint32_t *indices_out);
/* sum over queries of find().count(), nothing stored (the bench loop of benches/lapper_benchmark.rs:103-108) */
uint64_t lo_find_count_sum(const lo_lapper *l, const lo_coord *q_start, const lo_coord *q_stop, size_t nq,
                           int nthreads);

#ifdef __cplusplus
}
#endif
#endif
