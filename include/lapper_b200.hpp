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
//   to_bincode() / from_bincode()                 the `with_serde` feature     :85-86, :111-122, test :1387-1409
//
// `val: T` stays on the host; the engine returns positions into the sorted `intervals` vector.
// Errors: the reference panics; this facade throws lapper_b200::Error carrying the ABI status code.
#pragma once

#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <limits>
#include <stdexcept>
#include <string>
#include <type_traits>
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

// The C entry points behind each coordinate type (src/lib.rs:88-100: `I: PrimInt + Unsigned`): u32 -> lapper_*_u32,
// u64 (the README's usize) -> lapper64_*, u8 / u16 -> widened into the u32 entry points.
namespace detail {
template <typename I>
struct Abi;
template <>
struct Abi<uint32_t> {
    using handle = lapper_t;
    static int build(const uint32_t *s, const uint32_t *e, uint64_t n, int dev, uint32_t flags, handle **out) {
        return lapper_build_u32_ex(s, e, n, dev, flags, out);
    }
    static void release(handle *h) { lapper_free(h); }
    static uint64_t len(const handle *h) { return lapper_len(h); }
    static uint32_t max_len(const handle *h) { return lapper_max_len(h); }
    static bool overlaps_merged(const handle *h) { return lapper_overlaps_merged(h) != 0; }
    static int get_intervals(const handle *h, uint32_t *s, uint32_t *e, uint32_t *p) { return lapper_get_intervals(h, s, e, p); }
    static int count_batch(const handle *h, const uint32_t *a, const uint32_t *b, uint64_t nq, uint64_t *out) {
        return lapper_count_batch_u32(h, a, b, nq, out);
    }
    static int find_batch(const handle *h, const uint32_t *a, const uint32_t *b, uint64_t nq, lapper_result_t **out) {
        return lapper_find_batch_u32(h, a, b, nq, out);
    }
    static int seek_cursor(const handle *h, uint32_t start, uint64_t *c) { return lapper_seek_cursor(h, start, c); }
    static int insert(handle *h, uint32_t s, uint32_t e, uint64_t *pos) { return lapper_insert_u32(h, s, e, pos); }
    static int merge_overlaps(handle *h) { return lapper_merge_overlaps(h); }
    static int cov(const handle *h, uint32_t *c) { return lapper_cov(h, c); }
    static int set_cov(handle *h, uint32_t *c) { return lapper_set_cov(h, c); }
    static int union_and_intersect(const handle *a, const handle *b, uint32_t *u, uint32_t *i) {
        return lapper_union_and_intersect(a, b, u, i);
    }
    static int depth(const handle *h, uint64_t *n, uint32_t **s, uint32_t **e, uint32_t **d) { return lapper_depth(h, n, s, e, d); }
    static int tune_query_length(handle *h, uint32_t len) { return lapper_tune_query_length(h, len); }
    static int get_sorted_stops(const handle *h, uint32_t *out) { return lapper_get_stops(h, out); }
    static int get_cached_state(const handle *h, int *some, uint32_t *cov, int *merged) { return lapper_get_cached_state(h, some, cov, merged); }
    static int restore_cached_state(handle *h, int some, uint32_t cov, int merged) { return lapper_restore_cached_state(h, some, cov, merged); }
};
template <>
struct Abi<uint64_t> {
    using handle = lapper64_t;
    static int build(const uint64_t *s, const uint64_t *e, uint64_t n, int dev, uint32_t, handle **out) {
        return lapper64_build(s, e, n, dev, out);  // (the build flags shape the u32 tables only)
    }
    static void release(handle *h) { lapper64_free(h); }
    static uint64_t len(const handle *h) { return lapper64_len(h); }
    static uint64_t max_len(const handle *h) { return lapper64_max_len(h); }
    static bool overlaps_merged(const handle *h) { return lapper64_overlaps_merged(h) != 0; }
    static int get_intervals(const handle *h, uint64_t *s, uint64_t *e, uint32_t *p) { return lapper64_get_intervals(h, s, e, p); }
    static int count_batch(const handle *h, const uint64_t *a, const uint64_t *b, uint64_t nq, uint64_t *out) {
        return lapper64_count_batch(h, a, b, nq, out);
    }
    static int find_batch(const handle *h, const uint64_t *a, const uint64_t *b, uint64_t nq, lapper_result_t **out) {
        return lapper64_find_batch(h, a, b, nq, out);
    }
    static int seek_cursor(const handle *h, uint64_t start, uint64_t *c) { return lapper64_seek_cursor(h, start, c); }
    static int insert(handle *h, uint64_t s, uint64_t e, uint64_t *pos) { return lapper64_insert(h, s, e, pos); }
    static int merge_overlaps(handle *h) { return lapper64_merge_overlaps(h); }
    static int cov(const handle *h, uint64_t *c) { return lapper64_cov(h, c); }
    static int set_cov(handle *h, uint64_t *c) { return lapper64_set_cov(h, c); }
    static int union_and_intersect(const handle *a, const handle *b, uint64_t *u, uint64_t *i) {
        return lapper64_union_and_intersect(a, b, u, i);
    }
    static int depth(const handle *h, uint64_t *n, uint64_t **s, uint64_t **e, uint32_t **d) { return lapper64_depth(h, n, s, e, d); }
    static int tune_query_length(handle *h, uint32_t len) { return lapper64_tune_query_length(h, len); }
    static int get_sorted_stops(const handle *h, uint64_t *out) { return lapper64_get_sorted_stops(h, out); }
    static int get_cached_state(const handle *h, int *some, uint64_t *cov, int *merged) { return lapper64_get_cached_state(h, some, cov, merged); }
    static int restore_cached_state(handle *h, int some, uint64_t cov, int merged) { return lapper64_restore_cached_state(h, some, cov, merged); }
};
// Coordinate types narrower than 32 bits (u8, u16) widen losslessly into the u32 engine.  The one place where the width
// of I shows in the reference's arithmetic is `start + 1` in count (src/lib.rs:614), which wraps to 0 for start = I::MAX
// (release builds): the engine does the same for u32::MAX, so that is what such a start is widened to.  (find has no
// hit for it either way: no stop exceeds I::MAX.)  Results narrow back without loss: every coordinate, max_len and cov
// of a valid set fit I.
template <typename I>
struct WidenAbi {
    using handle = lapper_t;
    using B = Abi<uint32_t>;
    static constexpr I kMax = std::numeric_limits<I>::max();
    static std::vector<uint32_t> widen(const I *x, uint64_t n, bool as_query_start = false) {
        std::vector<uint32_t> w(n);
        for (uint64_t i = 0; i < n; i++) w[i] = (as_query_start && x[i] == kMax) ? 0xFFFFFFFFu : (uint32_t)x[i];
        return w;
    }
    static int build(const I *s, const I *e, uint64_t n, int dev, uint32_t flags, handle **out) {
        return B::build(widen(s, n).data(), widen(e, n).data(), n, dev, flags, out);
    }
    static void release(handle *h) { B::release(h); }
    static uint64_t len(const handle *h) { return B::len(h); }
    static I max_len(const handle *h) { return (I)B::max_len(h); }
    static bool overlaps_merged(const handle *h) { return B::overlaps_merged(h); }
    static int get_intervals(const handle *h, I *s, I *e, uint32_t *p) {
        const uint64_t n = B::len(h);
        std::vector<uint32_t> ws(n), we(n);
        const int rc = B::get_intervals(h, ws.data(), we.data(), p);
        for (uint64_t i = 0; i < n; i++) s[i] = (I)ws[i], e[i] = (I)we[i];
        return rc;
    }
    static int count_batch(const handle *h, const I *a, const I *b, uint64_t nq, uint64_t *out) {
        return B::count_batch(h, widen(a, nq, true).data(), widen(b, nq).data(), nq, out);
    }
    static int find_batch(const handle *h, const I *a, const I *b, uint64_t nq, lapper_result_t **out) {
        return B::find_batch(h, widen(a, nq, true).data(), widen(b, nq).data(), nq, out);
    }
    static int seek_cursor(const handle *h, I start, uint64_t *c) { return B::seek_cursor(h, (uint32_t)start, c); }
    static int insert(handle *h, I s, I e, uint64_t *pos) { return B::insert(h, (uint32_t)s, (uint32_t)e, pos); }
    static int merge_overlaps(handle *h) { return B::merge_overlaps(h); }
    static int cov(const handle *h, I *c) {
        uint32_t w = 0;
        const int rc = B::cov(h, &w);
        if (c) *c = (I)w;
        return rc;
    }
    static int set_cov(handle *h, I *c) {
        uint32_t w = 0;
        const int rc = B::set_cov(h, &w);
        if (c) *c = (I)w;
        return rc;
    }
    static int union_and_intersect(const handle *a, const handle *b, I *u, I *i) {
        uint32_t wu = 0, wi = 0;
        const int rc = B::union_and_intersect(a, b, &wu, &wi);
        *u = (I)wu, *i = (I)wi;
        return rc;
    }
    // the library's u32 arrays narrowed in place (front to back: element i is read before bytes it shares are written)
    static int depth(const handle *h, uint64_t *n, I **s, I **e, uint32_t **d) {
        uint32_t *ws = nullptr, *we = nullptr;
        const int rc = B::depth(h, n, &ws, &we, d);
        if (rc != LAPPER_OK) return rc;
        for (uint32_t *arr : {ws, we}) {
            unsigned char *bytes = reinterpret_cast<unsigned char *>(arr);  // (byte accesses: no aliasing assumptions)
            for (uint64_t i = 0; i < *n; i++) {
                uint32_t w;
                std::memcpy(&w, bytes + 4 * i, 4);
                const I v = (I)w;
                std::memcpy(bytes + sizeof(I) * i, &v, sizeof(I));
            }
        }
        *s = reinterpret_cast<I *>(ws), *e = reinterpret_cast<I *>(we);
        return rc;
    }
    static int tune_query_length(handle *h, uint32_t len) { return B::tune_query_length(h, len); }
    static int get_sorted_stops(const handle *h, I *out) {
        const uint64_t n = B::len(h);
        std::vector<uint32_t> w(n);
        const int rc = B::get_sorted_stops(h, w.data());
        for (uint64_t i = 0; i < n; i++) out[i] = (I)w[i];
        return rc;
    }
    static int get_cached_state(const handle *h, int *some, I *cov, int *merged) {
        uint32_t w = 0;
        const int rc = B::get_cached_state(h, some, &w, merged);
        if (cov) *cov = (I)w;
        return rc;
    }
    static int restore_cached_state(handle *h, int some, I cov, int merged) { return B::restore_cached_state(h, some, (uint32_t)cov, merged); }
};
template <>
struct Abi<uint16_t> : WidenAbi<uint16_t> {};
template <>
struct Abi<uint8_t> : WidenAbi<uint8_t> {};
}  // namespace detail

// src/lib.rs:88-100
// bincode 1.3 with default options (Cargo.lock pins 1.3.3): little-endian fixed-width integers, `Vec<X>` = u64 length
// + elements, `usize` = u64, `Option<X>` = one tag byte + X, `bool` = one byte, structs = their fields in order.
// (Host code; a little-endian host is assumed, as everywhere CUDA runs.)
namespace bincode {
template <typename X>
inline void put(std::vector<uint8_t> &out, X x) {
    static_assert(std::is_arithmetic<X>::value, "bincode: integer, bool or float fields only");
    uint8_t b[sizeof(X)];
    std::memcpy(b, &x, sizeof(X));
    out.insert(out.end(), b, b + sizeof(X));
}
template <>
inline void put<bool>(std::vector<uint8_t> &out, bool x) { out.push_back(x ? 1 : 0); }
struct Reader {
    const uint8_t *p, *end;
    template <typename X>
    X take() {
        static_assert(std::is_arithmetic<X>::value, "bincode: integer, bool or float fields only");
        if ((size_t)(end - p) < sizeof(X)) throw Error(LAPPER_ERR_INVALID_ARG, "bincode: unexpected end of input");
        X x;
        std::memcpy(&x, p, sizeof(X));
        p += sizeof(X);
        return x;
    }
    uint8_t tag(const char *what) {  // Option tag / bool: 0 or 1, anything else is malformed
        const uint8_t t = take<uint8_t>();
        if (t > 1) throw Error(LAPPER_ERR_INVALID_ARG, what);
        return t;
    }
    // a Vec length, checked against the bytes that are left (elem_size bytes per element)
    uint64_t vec_len(size_t elem_size) {
        const uint64_t n = take<uint64_t>();
        if (n > (uint64_t)(end - p) / (elem_size ? elem_size : 1)) throw Error(LAPPER_ERR_INVALID_ARG, "bincode: unexpected end of input");
        return n;
    }
};
template <>
inline bool Reader::take<bool>() { return tag("bincode: invalid bool") != 0; }
}  // namespace bincode

template <typename T, typename I = uint32_t>
struct Interval {
    I start;
    I stop;
    T val;
    // src/lib.rs:130-136
    I intersect(const Interval &o) const {
        const I hi = stop < o.stop ? stop : o.stop, lo = start > o.start ? start : o.start;
        return hi >= lo ? hi - lo : 0;
    }
    // src/lib.rs:138-142
    bool overlap(I s, I e) const { return start < e && stop > s; }
    // src/lib.rs:171-180: equality ignores val
    bool operator==(const Interval &o) const { return start == o.start && stop == o.stop; }
};

struct Csr {
    std::vector<uint64_t> offsets;  // nq + 1
    std::vector<uint32_t> indices;  // positions in Lapper::intervals, iterator order per query
};

// Lapper<I, T> (src/lib.rs:102-123); the coordinate type is the SECOND parameter here so that Lapper<T> stays the
// u32 form the reference's own tests and benches use (benches/lapper_benchmark.rs:13)
template <typename T, typename I = uint32_t>
class Lapper {
    using A = detail::Abi<I>;

  public:
    using Iv = Interval<T, I>;
    std::vector<Iv> intervals;  // sorted by (start, stop), stable  (src/lib.rs:112)

    explicit Lapper(std::vector<Iv> ivs, int device = 0, uint32_t flags = LAPPER_BUILD_DEFAULT) {
        std::vector<I> s(ivs.size()), e(ivs.size());
        for (size_t i = 0; i < ivs.size(); i++) s[i] = ivs[i].start, e[i] = ivs[i].stop;
        check(A::build(s.data(), e.data(), ivs.size(), device, flags, &h_));
        input_ = std::move(ivs);
        refresh();
    }
    ~Lapper() {
        if (h_) A::release(h_);
    }
    Lapper(const Lapper &) = delete;
    Lapper &operator=(const Lapper &) = delete;
    Lapper(Lapper &&o) noexcept : intervals(std::move(o.intervals)), h_(o.h_), input_(std::move(o.input_)) { o.h_ = nullptr; }

    size_t len() const { return intervals.size(); }       // src/lib.rs:284-287
    bool is_empty() const { return intervals.empty(); }   // src/lib.rs:296-299
    I max_len() const { return A::max_len(h_); }
    bool overlaps_merged() const { return A::overlaps_merged(h_); }
    const typename A::handle *handle() const { return h_; }
    typename std::vector<Iv>::const_iterator begin() const { return intervals.begin(); }  // iter(), :353-359
    typename std::vector<Iv>::const_iterator end() const { return intervals.end(); }

    // optional: announce the typical query length (stop - start) so that the packed count table's margin covers it
    void tune_query_length(uint32_t typical_query_len) { check(A::tune_query_length(h_, typical_query_len)); }

    // batch entry points
    std::vector<uint64_t> count_batch(const std::vector<I> &qs, const std::vector<I> &qe) const {
        if (qs.size() != qe.size()) throw Error(LAPPER_ERR_INVALID_ARG, "count_batch: q_start and q_stop differ in length");
        std::vector<uint64_t> out(qs.size());
        check(A::count_batch(h_, qs.data(), qe.data(), qs.size(), out.data()));
        return out;
    }
    Csr find_batch(const std::vector<I> &qs, const std::vector<I> &qe) const {
        if (qs.size() != qe.size()) throw Error(LAPPER_ERR_INVALID_ARG, "find_batch: q_start and q_stop differ in length");
        lapper_result_t *r = nullptr;
        check(A::find_batch(h_, qs.data(), qe.data(), qs.size(), &r));
        Csr c;
        c.offsets.resize(qs.size() + 1);
        c.indices.resize(lapper_result_total(r));
        const int rc = lapper_result_copy(r, c.offsets.data(), c.indices.data());
        lapper_result_free(r);
        check(rc);
        return c;
    }

    // per-query API of the reference
    uint64_t count(I start, I stop) const { return count_batch({start}, {stop})[0]; }
    std::vector<const Iv *> find(I start, I stop) const {
        std::vector<const Iv *> out;
        for (uint32_t i : find_batch({start}, {stop}).indices) out.push_back(&intervals[i]);
        return out;
    }
    std::vector<const Iv *> seek(I start, I stop, size_t *cursor) const {
        uint64_t c = *cursor;
        check(A::seek_cursor(h_, start, &c));  // the reference's cursor update, src/lib.rs:658-671
        *cursor = (size_t)c;
        std::vector<const Iv *> out;
        for (uint32_t i : find_batch({start}, {stop}).indices)
            if (i >= c) out.push_back(&intervals[i]);  // IterFind starts at the cursor
        return out;
    }

    // src/lib.rs:260-273 (O(n), like the reference)
    void insert(const Iv &elem) {
        uint64_t pos = 0;
        check(A::insert(h_, elem.start, elem.stop, &pos));
        if (input_.size() <= pos) input_.resize(pos + 1, elem);
        input_[pos] = elem;
        refresh();
    }
    void merge_overlaps() {
        check(A::merge_overlaps(h_));
        refresh();
    }
    I cov() const {
        I c = 0;
        check(A::cov(h_, &c));
        return c;
    }
    I set_cov() {
        I c = 0;
        check(A::set_cov(h_, &c));
        return c;
    }
    std::pair<I, I> union_and_intersect(const Lapper &other) const {
        I u = 0, i = 0;
        check(A::union_and_intersect(h_, other.h_, &u, &i));
        return {u, i};
    }
    // depth() / IterDepth (src/lib.rs:576-598, :720-784): runs of equal coverage depth, val = the depth
    std::vector<Interval<uint32_t, I>> depth() const {
        uint64_t n = 0;
        I *s = nullptr, *e = nullptr;
        uint32_t *d = nullptr;
        check(A::depth(h_, &n, &s, &e, &d));
        std::vector<Interval<uint32_t, I>> out;
        for (uint64_t i = 0; i < n; i++) out.push_back(Interval<uint32_t, I>{s[i], e[i], d[i]});
        lapper_host_free(s);
        lapper_host_free(e);
        lapper_host_free(d);
        return out;
    }
    I intersect(const Lapper &o) const { return union_and_intersect(o).second; }
    I union_(const Lapper &o) const { return union_and_intersect(o).first; }

    // ---- the `with_serde` feature (src/lib.rs:85-86): bincode::serialize(&lapper) / bincode::deserialize, fields in
    // the order of src/lib.rs:111-122.  T: an integer, bool or float type.
    std::vector<uint8_t> to_bincode() const {
        const size_t n = intervals.size();
        std::vector<I> stops(n);
        check(A::get_sorted_stops(h_, stops.data()));
        int some = 0, merged = 0;
        I cv = 0;
        check(A::get_cached_state(h_, &some, &cv, &merged));
        std::vector<uint8_t> out;
        out.reserve(8 + n * (2 * sizeof(I) + sizeof(T)) + 2 * (8 + n * sizeof(I)) + 2 * sizeof(I) + 2);
        bincode::put<uint64_t>(out, n);  // intervals: Vec<Interval<I, T>>
        for (const Iv &iv : intervals) bincode::put<I>(out, iv.start), bincode::put<I>(out, iv.stop), bincode::put<T>(out, iv.val);
        bincode::put<uint64_t>(out, n);  // starts: Vec<I>
        for (const Iv &iv : intervals) bincode::put<I>(out, iv.start);
        bincode::put<uint64_t>(out, n);  // stops: Vec<I> (sorted on their own)
        for (const I e : stops) bincode::put<I>(out, e);
        bincode::put<I>(out, max_len());
        bincode::put<uint8_t>(out, some ? 1 : 0);  // cov: Option<I>
        if (some) bincode::put<I>(out, cv);
        bincode::put<bool>(out, merged != 0);
        return out;
    }
    // The index is derived data: it is rebuilt from `intervals` and the serialised starts / stops / max_len are checked
    // against it; cov and overlaps_merged are restored as they were.  Malformed input throws Error(INVALID_ARG).
    static Lapper from_bincode(const uint8_t *data, size_t len, int device = 0) {
        bincode::Reader r{data, data + len};
        const uint64_t n = r.vec_len(2 * sizeof(I) + sizeof(T));
        std::vector<Iv> ivs;
        ivs.reserve(n);
        for (uint64_t i = 0; i < n; i++) {
            const I s = r.template take<I>(), e = r.template take<I>();
            ivs.push_back(Iv{s, e, r.template take<T>()});
        }
        std::vector<I> starts(r.vec_len(sizeof(I)));
        for (I &x : starts) x = r.template take<I>();
        std::vector<I> stops(r.vec_len(sizeof(I)));
        for (I &x : stops) x = r.template take<I>();
        const I max_len = r.template take<I>();
        const bool some = r.tag("bincode: invalid Option tag") != 0;
        const I cv = some ? r.template take<I>() : I(0);
        const bool merged = r.template take<bool>();
        if (r.p != r.end) throw Error(LAPPER_ERR_INVALID_ARG, "bincode: trailing bytes");
        Lapper lap(std::move(ivs), device);
        std::vector<I> have(lap.intervals.size());
        check(A::get_sorted_stops(lap.h_, have.data()));
        bool same = starts.size() == lap.intervals.size() && have == stops && lap.max_len() == max_len;
        for (size_t i = 0; same && i < starts.size(); i++) same = starts[i] == lap.intervals[i].start;
        if (!same) throw Error(LAPPER_ERR_INVALID_ARG, "bincode: serialised starts / stops / max_len do not belong to the serialised intervals");
        check(A::restore_cached_state(lap.h_, some ? 1 : 0, cv, merged ? 1 : 0));
        return lap;
    }
    static Lapper from_bincode(const std::vector<uint8_t> &buf, int device = 0) { return from_bincode(buf.data(), buf.size(), device); }

  private:
    // rebuild `intervals` from the engine's sorted order; vals follow through perm
    void refresh() {
        const size_t n = A::len(h_);
        std::vector<I> s(n), e(n);
        std::vector<uint32_t> p(n);
        check(A::get_intervals(h_, s.data(), e.data(), p.data()));
        intervals.clear();
        intervals.reserve(n);
        for (size_t i = 0; i < n; i++) intervals.push_back(Iv{s[i], e[i], input_[p[i]].val});
    }
    typename A::handle *h_ = nullptr;
    std::vector<Iv> input_;
};

}  // namespace lapper_b200
