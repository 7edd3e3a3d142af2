/*
 * lapper_oracle.c -- CPU restatement of rust-lapper 1.3.0.  TEST INFRASTRUCTURE ONLY
 * (see lapper_oracle.h).  Same AoS layout and the same loops as src/lib.rs so it is
 * also an honest single-thread CPU baseline; "C restatement, not rustc output".
 */
#include "lapper_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void *xmalloc(size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (!p) {
        fprintf(stderr, "lapper_oracle: out of memory (%zu bytes)\n", bytes);
        abort();
    }
    return p;
}

/* src/lib.rs:151-157  Ord for Interval: (start, stop); val ignored. */
static int iv_cmp(const lo_interval *a, const lo_interval *b) {
    if (a->start < b->start) return -1;
    if (a->start > b->start) return 1;
    if (a->stop < b->stop) return -1;
    if (a->stop > b->stop) return 1;
    return 0;
}

/* Vec::sort (src/lib.rs:201) is a stable sort; bottom-up merge sort keeps ties in input order. */
static void iv_stable_sort(lo_interval *a, size_t n) {
    if (n < 2) return;
    lo_interval *tmp = (lo_interval *)xmalloc(n * sizeof(lo_interval));
    lo_interval *src = a, *dst = tmp;
    for (size_t w = 1; w < n; w *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * w) {
            size_t mid = lo + w < n ? lo + w : n;
            size_t hi = lo + 2 * w < n ? lo + 2 * w : n;
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                /* take right only when strictly smaller: stability */
                if (iv_cmp(&src[j], &src[i]) < 0)
                    dst[k++] = src[j++];
                else
                    dst[k++] = src[i++];
            }
            while (i < mid) dst[k++] = src[i++];
            while (j < hi) dst[k++] = src[j++];
        }
        lo_interval *t = src;
        src = dst;
        dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(lo_interval));
    free(tmp);
}

static int u32_cmp(const void *a, const void *b) {
    lo_coord x = *(const lo_coord *)a, y = *(const lo_coord *)b;
    return (x > y) - (x < y);
}

/* src/lib.rs:203-227 (and :393-415): unzip, sort starts and stops, max_len with checked_sub -> 0 */
static void rebuild_aux(lo_lapper *l) {
    lo_coord max_len = 0;
    for (size_t i = 0; i < l->len; i++) {
        l->starts[i] = l->intervals[i].start;
        l->stops[i] = l->intervals[i].stop;
    }
    qsort(l->starts, l->len, sizeof(lo_coord), u32_cmp);
    qsort(l->stops, l->len, sizeof(lo_coord), u32_cmp);
    for (size_t i = 0; i < l->len; i++) {
        const lo_interval *iv = &l->intervals[i];
        lo_coord i_len = iv->stop >= iv->start ? iv->stop - iv->start : 0; /* checked_sub.unwrap_or(0) */
        if (i_len > max_len) max_len = i_len;
    }
    l->max_len = max_len;
}

lo_lapper *lo_new(const lo_coord *starts, const lo_coord *stops, const uint32_t *vals, size_t n) {
    lo_lapper *l = (lo_lapper *)xmalloc(sizeof(lo_lapper));
    l->len = n;
    l->cap = n ? n : 1;
    l->intervals = (lo_interval *)xmalloc(l->cap * sizeof(lo_interval));
    l->starts = (lo_coord *)xmalloc(l->cap * sizeof(lo_coord));
    l->stops = (lo_coord *)xmalloc(l->cap * sizeof(lo_coord));
    for (size_t i = 0; i < n; i++) {
        l->intervals[i].start = starts[i];
        l->intervals[i].stop = stops[i];
        l->intervals[i].val = vals ? vals[i] : (uint32_t)i;
    }
    iv_stable_sort(l->intervals, n); /* :201 */
    rebuild_aux(l);                  /* :203-227 */
    l->cov_is_some = 0;              /* :233 */
    l->cov = 0;
    l->overlaps_merged = 0; /* :234 */
    return l;
}

void lo_free(lo_lapper *l) {
    if (!l) return;
    free(l->intervals);
    free(l->starts);
    free(l->stops);
    free(l);
}

lo_coord lo_interval_intersect(lo_coord s1, lo_coord e1, lo_coord s2, lo_coord e2) {
    /* :133-135  min(stops).checked_sub(max(starts)).unwrap_or(0) */
    lo_coord mn = e1 < e2 ? e1 : e2;
    lo_coord mx = s1 > s2 ? s1 : s2;
    return mn >= mx ? mn - mx : 0;
}

int lo_interval_overlap(const lo_interval *iv, lo_coord start, lo_coord stop) {
    return iv->start < stop && iv->stop > start; /* :141 */
}

size_t lo_lower_bound(lo_coord start, const lo_interval *intervals, size_t n) {
    /* :424-436 */
    size_t size = n;
    size_t low = 0;
    while (size > 0) {
        size_t half = size / 2;
        size_t other_half = size - half;
        size_t probe = low + half;
        size_t other_low = low + other_half;
        const lo_interval *v = &intervals[probe];
        size = half;
        low = v->start < start ? other_low : low;
    }
    return low;
}

size_t lo_bsearch_seq(lo_coord key, const lo_coord *elems, size_t n) {
    /* :452-465 */
    if (n == 0 || elems[0] >= key) {
        return 0;
    } else if (elems[n - 1] < key) {
        return n;
    }
    size_t cursor = 0;
    size_t length = n;
    while (length > 1) {
        size_t half = length >> 1;
        length -= half;
        cursor += (size_t)(elems[cursor + half - 1] < key) * half;
    }
    return cursor;
}

/* bsearch_seq_ref over the interval vector with Interval's PartialOrd (insert, :263) */
static size_t bsearch_seq_iv(const lo_interval *key, const lo_interval *elems, size_t n) {
    if (n == 0 || iv_cmp(&elems[0], key) >= 0) {
        return 0;
    } else if (iv_cmp(&elems[n - 1], key) < 0) {
        return n;
    }
    size_t cursor = 0;
    size_t length = n;
    while (length > 1) {
        size_t half = length >> 1;
        length -= half;
        cursor += (size_t)(iv_cmp(&elems[cursor + half - 1], key) < 0) * half;
    }
    return cursor;
}

void lo_insert(lo_lapper *l, lo_coord start, lo_coord stop, uint32_t val) {
    /* :261-272 */
    lo_interval elem = {start, stop, val};
    size_t starts_insert_index = lo_bsearch_seq(start, l->starts, l->len);
    size_t stops_insert_index = lo_bsearch_seq(stop, l->stops, l->len);
    size_t intervals_insert_index = bsearch_seq_iv(&elem, l->intervals, l->len);
    lo_coord i_len = stop >= start ? stop - start : 0;
    if (i_len > l->max_len) l->max_len = i_len;
    if (l->len + 1 > l->cap) {
        l->cap = l->cap * 2 + 1;
        l->intervals = (lo_interval *)realloc(l->intervals, l->cap * sizeof(lo_interval));
        l->starts = (lo_coord *)realloc(l->starts, l->cap * sizeof(lo_coord));
        l->stops = (lo_coord *)realloc(l->stops, l->cap * sizeof(lo_coord));
        if (!l->intervals || !l->starts || !l->stops) abort();
    }
    memmove(&l->starts[starts_insert_index + 1], &l->starts[starts_insert_index],
            (l->len - starts_insert_index) * sizeof(lo_coord));
    l->starts[starts_insert_index] = start;
    memmove(&l->stops[stops_insert_index + 1], &l->stops[stops_insert_index],
            (l->len - stops_insert_index) * sizeof(lo_coord));
    l->stops[stops_insert_index] = stop;
    memmove(&l->intervals[intervals_insert_index + 1], &l->intervals[intervals_insert_index],
            (l->len - intervals_insert_index) * sizeof(lo_interval));
    l->intervals[intervals_insert_index] = elem;
    l->len += 1;
    l->cov_is_some = 0;
    l->overlaps_merged = 0;
}

uint64_t lo_count(const lo_lapper *l, lo_coord start, lo_coord stop) {
    /* :612-617; `start + one` wraps like a release build when start == u32::MAX */
    uint64_t len = (uint64_t)l->len;
    uint64_t first = (uint64_t)lo_bsearch_seq((lo_coord)(start + 1u), l->stops, l->len);
    uint64_t last = (uint64_t)lo_bsearch_seq(stop, l->starts, l->len);
    uint64_t num_cant_after = len - last;
    return len - first - num_cant_after; /* wrapping usize */
}

lo_iter_find lo_find(const lo_lapper *l, lo_coord start, lo_coord stop) {
    /* :630-638 */
    lo_iter_find it;
    it.inner = l;
    it.off = lo_lower_bound(start >= l->max_len ? start - l->max_len : 0, l->intervals, l->len);
    it.start = start;
    it.stop = stop;
    return it;
}

lo_iter_find lo_seek(const lo_lapper *l, lo_coord start, lo_coord stop, size_t *cursor) {
    /* :658-678 */
    lo_coord floor = start >= l->max_len ? start - l->max_len : 0;
    if (*cursor == 0 || (*cursor < l->len && l->intervals[*cursor].start > start)) {
        *cursor = lo_lower_bound(floor, l->intervals, l->len);
    }
    while (*cursor + 1 < l->len && l->intervals[*cursor + 1].start < floor) {
        *cursor += 1;
    }
    lo_iter_find it;
    it.inner = l;
    it.off = *cursor;
    it.start = start;
    it.stop = stop;
    return it;
}

int64_t lo_iter_find_next(lo_iter_find *it) {
    /* :705-716 */
    while (it->off < it->inner->len) {
        const lo_interval *interval = &it->inner->intervals[it->off];
        it->off += 1;
        if (lo_interval_overlap(interval, it->start, it->stop)) {
            return (int64_t)(it->off - 1);
        } else if (interval->start >= it->stop) {
            break;
        }
    }
    return -1;
}

lo_coord lo_calculate_coverage(const lo_lapper *l) {
    /* :328-349; u32 arithmetic wraps like a release build */
    lo_interval moving = {0, 0, 0};
    lo_coord cov = 0;
    for (size_t i = 0; i < l->len; i++) {
        const lo_interval *interval = &l->intervals[i];
        if (lo_interval_overlap(&moving, interval->start, interval->stop)) {
            moving.start = moving.start < interval->start ? moving.start : interval->start;
            moving.stop = moving.stop > interval->stop ? moving.stop : interval->stop;
        } else {
            cov = cov + (moving.stop - moving.start);
            moving.start = interval->start;
            moving.stop = interval->stop;
        }
    }
    cov = cov + (moving.stop - moving.start);
    return cov;
}

lo_coord lo_cov(const lo_lapper *l) {
    return l->cov_is_some ? l->cov : lo_calculate_coverage(l); /* :311-316 */
}

lo_coord lo_set_cov(lo_lapper *l) {
    lo_coord cov = lo_calculate_coverage(l); /* :320-324 */
    l->cov_is_some = 1;
    l->cov = cov;
    return cov;
}

void lo_merge_overlaps(lo_lapper *l) {
    /* :364-415.  The VecDeque of &mut Interval is a stack whose top is mutated in place;
     * writing the stack into the front of the same array is the same thing. */
    if (l->len > 0) {
        size_t top = 0; /* stack = intervals[0..=top] */
        for (size_t i = 1; i < l->len; i++) {
            const lo_interval interval = l->intervals[i];
            lo_interval *t = &l->intervals[top];
            if (t->stop < interval.start) {
                top += 1;
                l->intervals[top] = interval;
            } else if (t->stop < interval.stop) {
                t->stop = interval.stop;
            } else {
                /* they were equal (or contained) */
            }
        }
        l->overlaps_merged = 1;
        l->len = top + 1;
    }
    rebuild_aux(l); /* :393-415 */
}

void lo_union_and_intersect(const lo_lapper *self, const lo_lapper *other, lo_coord *union_out,
                            lo_coord *intersect_out) {
    /* :514-544 */
    size_t cursor = 0;
    if (!self->overlaps_merged || !other->overlaps_merged) {
        size_t cap = 16, cnt = 0;
        lo_coord *is = (lo_coord *)xmalloc(cap * sizeof(lo_coord));
        lo_coord *ie = (lo_coord *)xmalloc(cap * sizeof(lo_coord));
        for (size_t i = 0; i < self->len; i++) {
            const lo_interval *self_iv = &self->intervals[i];
            lo_iter_find it = lo_seek(other, self_iv->start, self_iv->stop, &cursor);
            int64_t j;
            while ((j = lo_iter_find_next(&it)) >= 0) {
                const lo_interval *other_iv = &other->intervals[j];
                if (cnt == cap) {
                    cap *= 2;
                    is = (lo_coord *)realloc(is, cap * sizeof(lo_coord));
                    ie = (lo_coord *)realloc(ie, cap * sizeof(lo_coord));
                    if (!is || !ie) abort();
                }
                is[cnt] = self_iv->start > other_iv->start ? self_iv->start : other_iv->start;
                ie[cnt] = self_iv->stop < other_iv->stop ? self_iv->stop : other_iv->stop;
                cnt++;
            }
        }
        lo_lapper *temp = lo_new(is, ie, NULL, cnt);
        lo_merge_overlaps(temp);
        lo_set_cov(temp);
        lo_coord u = lo_cov(self) + lo_cov(other) - lo_cov(temp);
        *union_out = u;
        *intersect_out = lo_cov(temp);
        lo_free(temp);
        free(is);
        free(ie);
    } else {
        lo_coord intersect = 0;
        for (size_t i = 0; i < self->len; i++) {
            const lo_interval *c1 = &self->intervals[i];
            lo_iter_find it = lo_seek(other, c1->start, c1->stop, &cursor);
            int64_t j;
            while ((j = lo_iter_find_next(&it)) >= 0) {
                const lo_interval *c2 = &other->intervals[j];
                intersect = intersect + lo_interval_intersect(c1->start, c1->stop, c2->start, c2->stop);
            }
        }
        *union_out = lo_cov(self) + lo_cov(other) - intersect;
        *intersect_out = intersect;
    }
}

lo_iter_depth lo_depth(const lo_lapper *l) {
    /* :578-597 */
    lo_coord *s = (lo_coord *)xmalloc(l->len * sizeof(lo_coord));
    lo_coord *e = (lo_coord *)xmalloc(l->len * sizeof(lo_coord));
    for (size_t i = 0; i < l->len; i++) {
        s[i] = l->intervals[i].start;
        e[i] = l->intervals[i].stop;
    }
    lo_lapper *merged = lo_new(s, e, NULL, l->len);
    free(s);
    free(e);
    lo_merge_overlaps(merged);
    lo_iter_depth it;
    it.inner = l;
    it.merged = merged;
    it.curr_merged_pos = 0;
    it.curr_pos = 0;
    it.cursor = 0;
    it.end = merged->len;
    return it;
}

static size_t seek_count(const lo_lapper *l, lo_coord start, lo_coord stop, size_t *cursor) {
    lo_iter_find it = lo_seek(l, start, stop, cursor);
    size_t c = 0;
    while (lo_iter_find_next(&it) >= 0) c++;
    return c;
}

int lo_iter_depth_next(lo_iter_depth *it, lo_coord *start_out, lo_coord *stop_out, uint32_t *val_out) {
    /* :743-783 */
    if (it->curr_pos >= it->merged->len) {
        fprintf(stderr, "lapper_oracle: depth() index out of bounds (the reference panics, src/lib.rs:744)\n");
        abort();
    }
    const lo_interval *interval = &it->merged->intervals[it->curr_pos];
    if (it->curr_merged_pos == 0) {
        it->curr_merged_pos = interval->start;
    }
    if (interval->stop == it->curr_merged_pos) {
        if (it->curr_pos + 1 != it->end) {
            it->curr_pos += 1;
            interval = &it->merged->intervals[it->curr_pos];
            it->curr_merged_pos = interval->start;
        } else {
            return 0;
        }
    }
    lo_coord start = it->curr_merged_pos;
    size_t depth_at_point = seek_count(it->inner, it->curr_merged_pos, it->curr_merged_pos + 1u, &it->cursor);
    size_t new_depth_at_point = depth_at_point;
    while (new_depth_at_point == depth_at_point && it->curr_merged_pos < interval->stop) {
        it->curr_merged_pos = it->curr_merged_pos + 1u;
        new_depth_at_point = seek_count(it->inner, it->curr_merged_pos, it->curr_merged_pos + 1u, &it->cursor);
    }
    *start_out = start;
    *stop_out = it->curr_merged_pos;
    *val_out = (uint32_t)depth_at_point;
    return 1;
}

void lo_iter_depth_free(lo_iter_depth *it) {
    lo_free(it->merged);
    it->merged = NULL;
}

/* ------------------------------------------------------------------ batch drivers */

typedef struct {
    const lo_lapper *l;
    const lo_coord *qs, *qe;
    size_t lo, hi;
    int mode; /* 0 count, 1 find-count into counts, 2 find-fill, 3 find-count sum */
    uint64_t *counts;
    const uint64_t *offsets;
    uint32_t *indices;
    uint64_t sum;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    const lo_lapper *l = j->l;
    uint64_t sum = 0;
    for (size_t q = j->lo; q < j->hi; q++) {
        if (j->mode == 0) {
            j->counts[q] = lo_count(l, j->qs[q], j->qe[q]);
        } else {
            lo_iter_find it = lo_find(l, j->qs[q], j->qe[q]);
            int64_t pos;
            uint64_t c = 0;
            if (j->mode == 2) {
                uint32_t *out = j->indices + j->offsets[q];
                while ((pos = lo_iter_find_next(&it)) >= 0) out[c++] = (uint32_t)pos;
            } else {
                while ((pos = lo_iter_find_next(&it)) >= 0) c++;
                if (j->mode == 1) j->counts[q] = c;
            }
            sum += c;
        }
    }
    j->sum = sum;
    return NULL;
}

static uint64_t run_batch(batch_job proto, size_t nq, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > nq) nthreads = nq ? (int)nq : 1;
    batch_job *jobs = (batch_job *)xmalloc((size_t)nthreads * sizeof(batch_job));
    pthread_t *tids = (pthread_t *)xmalloc((size_t)nthreads * sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = proto;
        jobs[t].lo = nq * (size_t)t / (size_t)nthreads;
        jobs[t].hi = nq * (size_t)(t + 1) / (size_t)nthreads;
        jobs[t].sum = 0;
    }
    if (nthreads == 1) {
        batch_worker(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; t++) pthread_create(&tids[t], NULL, batch_worker, &jobs[t]);
        for (int t = 0; t < nthreads; t++) pthread_join(tids[t], NULL);
    }
    uint64_t sum = 0;
    for (int t = 0; t < nthreads; t++) sum += jobs[t].sum;
    free(jobs);
    free(tids);
    return sum;
}

void lo_count_batch(const lo_lapper *l, const lo_coord *q_start, const lo_coord *q_stop, size_t nq,
                    uint64_t *counts_out, int nthreads) {
    batch_job p;
    memset(&p, 0, sizeof(p));
    p.l = l;
    p.qs = q_start;
    p.qe = q_stop;
    p.mode = 0;
    p.counts = counts_out;
    run_batch(p, nq, nthreads);
}

uint64_t lo_find_batch_offsets(const lo_lapper *l, const lo_coord *q_start, const lo_coord *q_stop,
                               size_t nq, uint64_t *offsets_out, int nthreads) {
    batch_job p;
    memset(&p, 0, sizeof(p));
    p.l = l;
    p.qs = q_start;
    p.qe = q_stop;
    p.mode = 1;
    p.counts = offsets_out; /* counts first, then exclusive scan in place */
    run_batch(p, nq, nthreads);
    uint64_t run = 0;
    for (size_t q = 0; q < nq; q++) {
        uint64_t c = offsets_out[q];
        offsets_out[q] = run;
        run += c;
    }
    offsets_out[nq] = run;
    return run;
}

void lo_find_batch_fill(const lo_lapper *l, const lo_coord *q_start, const lo_coord *q_stop, size_t nq,
                        const uint64_t *offsets, uint32_t *indices_out, int nthreads) {
    batch_job p;
    memset(&p, 0, sizeof(p));
    p.l = l;
    p.qs = q_start;
    p.qe = q_stop;
    p.mode = 2;
    p.offsets = offsets;
    p.indices = indices_out;
    run_batch(p, nq, nthreads);
}

void lo_seek_batch_fill(const lo_lapper *l, const lo_coord *q_start, const lo_coord *q_stop, size_t nq,
                        const uint64_t *offsets, uint32_t *indices_out) {
    size_t cursor = 0;
    for (size_t q = 0; q < nq; q++) {
        lo_iter_find it = lo_seek(l, q_start[q], q_stop[q], &cursor);
        uint32_t *out = indices_out + offsets[q];
        int64_t pos;
        uint64_t c = 0;
        uint64_t room = offsets[q + 1] - offsets[q];
        while ((pos = lo_iter_find_next(&it)) >= 0) {
            if (c < room) out[c] = (uint32_t)pos;
            c++;
        }
    }
}

uint64_t lo_find_count_sum(const lo_lapper *l, const lo_coord *q_start, const lo_coord *q_stop, size_t nq,
                           int nthreads) {
    batch_job p;
    memset(&p, 0, sizeof(p));
    p.l = l;
    p.qs = q_start;
    p.qe = q_stop;
    p.mode = 3;
    return run_batch(p, nq, nthreads);
}
