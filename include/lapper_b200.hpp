// lapper_b200.hpp -- header-only C++ facade over the C ABI (lapper_b200.h), mirroring the reference's
// `Lapper<I, T>` with I = u32 (/root/reference/src/lib.rs:182-680): same method names, argument meaning and
// result order, so code written against rust-lapper ports line by line:
//
//   Lapper<T>::Lapper(std::vector<Interval<T>>)   Lapper::new             src/lib.rs:196-236
//   find(start, stop)                             Lapper::find            :628-639 (+ IterFind::next :695-718)
//   seek(start, stop, cursor)                     Lapper::seek            :656-679
//   count(start, stop)                            Lapper::count           :610-618
//   merge_overlaps(), cov(), set_cov()            :361-416, :310-324
//   union_and_intersect / intersect / union_      :512-559
//   count_batch / find_batch                      new batch entry points (CSR)
//
// `val: T` stays on the host; the engine returns positions into the sorted `intervals` vector.
// Errors: the reference panics; this facade throws lapper_b200::Error carrying the ABI status code.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "lapper_b200.h"

namespace lapper_b200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const char *msg) : std::runtime_error(std::string("liblapper_b200: ") + (msg ? msg : "")), code(c) {}
};
inline void check(int rc) {
    if (rc != LAPPER_OK) throw Error(rc, lapper_last_error());
}

// src/lib.rs:88-100
template <typename T>
struct Interval {
    uint32_t start;
    uint32_t stop;
    T val;
    // src/lib.rs:130-136
    uint32_t intersect(const Interval &o) const {
        const uint32_t hi = stop < o.stop ? stop : o.stop, lo = start > o.start ? start : o.start;
        return hi >= lo ? hi - lo : 0;
    }
    // src/lib.rs:138-142
    bool overlap(uint32_t s, uint32_t e) const { return start < e && stop > s; }
    // src/lib.rs:171-180: equality ignores val
    bool operator==(const Interval &o) const { return start == o.start && stop == o.stop; }
};

struct Csr {
    std::vector<uint64_t> offsets;  // nq + 1
    std::vector<uint32_t> indices;  // positions in Lapper::intervals, iterator order per query
};

template <typename T>
class Lapper {
  public:
    std::vector<Interval<T>> intervals;  // sorted by (start, stop), stable  (src/lib.rs:112)

    explicit Lapper(std::vector<Interval<T>> ivs, int device = 0, uint32_t flags = LAPPER_BUILD_DEFAULT) {
        std::vector<uint32_t> s(ivs.size()), e(ivs.size());
        for (size_t i = 0; i < ivs.size(); i++) s[i] = ivs[i].start, e[i] = ivs[i].stop;
        check(lapper_build_u32_ex(s.data(), e.data(), ivs.size(), device, flags, &h_));
        input_ = std::move(ivs);
        refresh();
    }
    ~Lapper() { lapper_free(h_); }
    Lapper(const Lapper &) = delete;
    Lapper &operator=(const Lapper &) = delete;

    size_t len() const { return intervals.size(); }       // src/lib.rs:284-287
    bool is_empty() const { return intervals.empty(); }   // src/lib.rs:296-299
    uint32_t max_len() const { return lapper_max_len(h_); }
    bool overlaps_merged() const { return lapper_overlaps_merged(h_) != 0; }
    const lapper_t *handle() const { return h_; }
    typename std::vector<Interval<T>>::const_iterator begin() const { return intervals.begin(); }  // iter(), :353-359
    typename std::vector<Interval<T>>::const_iterator end() const { return intervals.end(); }

    // batch entry points
    std::vector<uint64_t> count_batch(const std::vector<uint32_t> &qs, const std::vector<uint32_t> &qe) const {
        if (qs.size() != qe.size()) throw Error(LAPPER_ERR_INVALID_ARG, "count_batch: q_start and q_stop differ in length");
        std::vector<uint64_t> out(qs.size());
        check(lapper_count_batch_u32(h_, qs.data(), qe.data(), qs.size(), out.data()));
        return out;
    }
    Csr find_batch(const std::vector<uint32_t> &qs, const std::vector<uint32_t> &qe) const {
        if (qs.size() != qe.size()) throw Error(LAPPER_ERR_INVALID_ARG, "find_batch: q_start and q_stop differ in length");
        lapper_result_t *r = nullptr;
        check(lapper_find_batch_u32(h_, qs.data(), qe.data(), qs.size(), &r));
        Csr c;
        c.offsets.resize(qs.size() + 1);
        c.indices.resize(lapper_result_total(r));
        const int rc = lapper_result_copy(r, c.offsets.data(), c.indices.data());
        lapper_result_free(r);
        check(rc);
        return c;
    }

    // per-query API of the reference
    uint64_t count(uint32_t start, uint32_t stop) const { return count_batch({start}, {stop})[0]; }
    std::vector<const Interval<T> *> find(uint32_t start, uint32_t stop) const {
        std::vector<const Interval<T> *> out;
        for (uint32_t i : find_batch({start}, {stop}).indices) out.push_back(&intervals[i]);
        return out;
    }
    std::vector<const Interval<T> *> seek(uint32_t start, uint32_t stop, size_t *cursor) const {
        uint64_t c = *cursor;
        check(lapper_seek_cursor(h_, start, &c));  // the reference's cursor update, src/lib.rs:658-671
        *cursor = (size_t)c;
        std::vector<const Interval<T> *> out;
        for (uint32_t i : find_batch({start}, {stop}).indices)
            if (i >= c) out.push_back(&intervals[i]);  // IterFind starts at the cursor
        return out;
    }

    // src/lib.rs:260-273 (O(n), like the reference)
    void insert(const Interval<T> &elem) {
        uint64_t pos = 0;
        check(lapper_insert_u32(h_, elem.start, elem.stop, &pos));
        if (input_.size() <= pos) input_.resize(pos + 1, elem);
        input_[pos] = elem;
        refresh();
    }
    void merge_overlaps() {
        check(lapper_merge_overlaps(h_));
        refresh();
    }
    uint32_t cov() const {
        uint32_t c = 0;
        check(lapper_cov(h_, &c));
        return c;
    }
    uint32_t set_cov() {
        uint32_t c = 0;
        check(lapper_set_cov(h_, &c));
        return c;
    }
    std::pair<uint32_t, uint32_t> union_and_intersect(const Lapper &other) const {
        uint32_t u = 0, i = 0;
        check(lapper_union_and_intersect(h_, other.h_, &u, &i));
        return {u, i};
    }
    // depth() / IterDepth (src/lib.rs:576-598, :720-784): runs of equal coverage depth, val = the depth
    std::vector<Interval<uint32_t>> depth() const {
        uint64_t n = 0;
        uint32_t *s = nullptr, *e = nullptr, *d = nullptr;
        check(lapper_depth(h_, &n, &s, &e, &d));
        std::vector<Interval<uint32_t>> out;
        for (uint64_t i = 0; i < n; i++) out.push_back(Interval<uint32_t>{s[i], e[i], d[i]});
        lapper_host_free(s);
        lapper_host_free(e);
        lapper_host_free(d);
        return out;
    }
    uint32_t intersect(const Lapper &o) const { return union_and_intersect(o).second; }
    uint32_t union_(const Lapper &o) const { return union_and_intersect(o).first; }

  private:
    // rebuild `intervals` from the engine's sorted order; vals follow through perm
    void refresh() {
        const size_t n = lapper_len(h_);
        std::vector<uint32_t> s(n), e(n), p(n);
        check(lapper_get_intervals(h_, s.data(), e.data(), p.data()));
        intervals.clear();
        intervals.reserve(n);
        for (size_t i = 0; i < n; i++) intervals.push_back(Interval<T>{s[i], e[i], input_[p[i]].val});
    }
    lapper_t *h_ = nullptr;
    std::vector<Interval<T>> input_;
};

}  // namespace lapper_b200
