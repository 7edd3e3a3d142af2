// The reference's unit tests (/root/reference/src/lib.rs `mod tests`) restated against the C++ facade
// include/lapper_b200.hpp -- they read like the originals.  Needs a B200 to run; compiles anywhere.
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>

#include "lapper_b200.hpp"

using lapper_b200::Interval;

static int failures = 0;
#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) {                                                   \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                  \
        }                                                                \
    } while (0)

// The suite runs once per coordinate type: I = u32 (the reference's own tests and benches) and I = u64 (the README's
// usize; these coordinates fit 32 bits, so batches of >= 4096 queries would be narrowed -- single queries are not).
template <typename I>
struct Suite {
    typedef Interval<uint32_t, I> Iv;
    template <typename T>
    using Lapper = lapper_b200::Lapper<T, I>;

    static std::vector<Iv> step_by(I lo, I hi, I step, I len) {
        std::vector<Iv> v;
        for (I x = lo; x < hi; x += step) v.push_back(Iv{x, x + len, 0});
        return v;
    }
    static std::vector<Iv> nonoverlapping() { return step_by(0, 100, 20, 10); }  // :857-867
    static std::vector<Iv> overlapping() { return step_by(0, 100, 10, 15); }     // :869-879
    static std::vector<Iv> badlapper() {                                         // :881-895
        return {{70, 120, 0}, {10, 15, 1}, {10, 15, 2}, {12, 15, 3}, {14, 16, 4}, {40, 45, 5},
                {50, 55, 6},  {60, 65, 7}, {68, 71, 8}, {70, 75, 9}};
    }
    static bool same(const std::vector<const Iv *> &got, const std::vector<Iv> &want) {
        if (got.size() != want.size()) return false;
        for (size_t i = 0; i < got.size(); i++)
            if (!(*got[i] == want[i])) return false;
        return true;
    }
    static void run();
};

template <typename I>
void Suite<I>::run() {
    {  // test_query_stop_interval_start / test_query_start_interval_stop  :1023-1039
        Lapper<uint32_t> lapper(nonoverlapping());
        size_t cursor = 0;
        CHECK(lapper.find(15, 20).empty());
        CHECK(lapper.seek(15, 20, &cursor).empty());
        CHECK(lapper.find(15, 20).size() == lapper.count(15, 20));
        CHECK(lapper.find(30, 35).empty());
        CHECK(lapper.find(30, 35).size() == lapper.count(30, 35));
    }
    {  // the four overlap shapes  :1043-1099
        Lapper<uint32_t> lapper(nonoverlapping());
        const I qs[4][2] = {{15, 25}, {25, 35}, {22, 27}, {15, 35}};
        for (auto &q : qs) {
            size_t cursor = 0;
            const Iv expected{20, 30, 0};
            CHECK(*lapper.find(q[0], q[1]).front() == expected);
            CHECK(*lapper.seek(q[0], q[1], &cursor).front() == expected);
            CHECK(lapper.find(q[0], q[1]).size() == lapper.count(q[0], q[1]));
        }
    }
    {  // test_overlapping_intervals  :1102-1121
        Lapper<uint32_t> lapper(overlapping());
        size_t cursor = 0;
        CHECK(same(lapper.find(8, 20), {{0, 15, 0}, {10, 25, 0}}));
        CHECK(same(lapper.seek(8, 20, &cursor), {{0, 15, 0}, {10, 25, 0}}));
        CHECK(lapper.count(8, 20) == 2);
    }
    {  // test_merge_overlaps  :1124-1138  (+ val of a run is the val of its first interval, :385-389)
        Lapper<uint32_t> lapper(badlapper());
        lapper.merge_overlaps();
        const std::vector<Iv> expected = {{10, 16, 0}, {40, 45, 0}, {50, 55, 0}, {60, 65, 0}, {68, 120, 0}};
        CHECK(lapper.intervals.size() == expected.size());
        for (size_t i = 0; i < expected.size() && i < lapper.intervals.size(); i++) CHECK(lapper.intervals[i] == expected[i]);
        CHECK(lapper.intervals[0].val == 1 && lapper.intervals[4].val == 8);
        CHECK(lapper.overlaps_merged());
    }
    {  // test_merge_overlaps_find  :1143-1165
        Lapper<uint32_t> lapper({{2, 3, 0}, {3, 4, 0}, {4, 6, 0}, {6, 7, 0}, {7, 8, 0}});
        CHECK(same(lapper.find(7, 9), {{7, 8, 0}}));
        lapper.merge_overlaps();
        CHECK(same(lapper.find(7, 9), {{2, 8, 0}}));
    }
    {  // test_lapper_cov  :1168-1178
        Lapper<uint32_t> lapper(badlapper());
        const I before = lapper.cov();
        lapper.merge_overlaps();
        CHECK(before == lapper.cov() && before == 73);
        Lapper<uint32_t> l2(nonoverlapping());
        l2.set_cov();
        CHECK(l2.cov() == 50);
    }
    {  // test_interval_intersects  :1181-1200
        const Iv i1{70, 120, 0}, i2{10, 15, 0}, i3{10, 15, 0}, i4{12, 15, 0}, i5{14, 16, 0}, i6{40, 50, 0}, i7{50, 55, 0},
            i8{60, 65, 0}, i9{68, 71, 0}, i10{70, 75, 0};
        CHECK(i2.intersect(i3) == 5 && i2.intersect(i4) == 3 && i2.intersect(i5) == 1 && i9.intersect(i10) == 1);
        CHECK(i7.intersect(i8) == 0 && i6.intersect(i7) == 0 && i1.intersect(i10) == 5);
    }
    {  // test_union_and_intersect  :1203-1240
        Lapper<uint32_t> l1({{70, 120, 0}, {10, 15, 0}, {12, 15, 0}, {14, 16, 0}, {68, 71, 0}});
        Lapper<uint32_t> l2({{10, 15, 0}, {40, 45, 0}, {50, 55, 0}, {60, 65, 0}, {70, 75, 0}});
        CHECK(l1.union_and_intersect(l2) == std::make_pair((I)73, (I)10));
        CHECK(l2.union_and_intersect(l1) == std::make_pair((I)73, (I)10));
        l1.merge_overlaps();
        l1.set_cov();
        l2.merge_overlaps();
        l2.set_cov();
        CHECK(l1.union_and_intersect(l2) == std::make_pair((I)73, (I)10));
        CHECK(l2.union_and_intersect(l1) == std::make_pair((I)73, (I)10));
    }
    {  // test_find_overlaps_in_large_intervals  :1243-1276
        Lapper<uint32_t> lapper({{0, 8, 0}, {1, 10, 0}, {2, 5, 0}, {3, 8, 0}, {4, 7, 0}, {5, 8, 0}, {8, 8, 0}, {9, 11, 0},
                                 {10, 13, 0}, {100, 200, 0}, {110, 120, 0}, {110, 124, 0}, {111, 160, 0}, {150, 200, 0}});
        CHECK(same(lapper.find(8, 11), {{1, 10, 0}, {9, 11, 0}, {10, 13, 0}}));
        CHECK(lapper.count(8, 11) == 3);
        CHECK(same(lapper.find(145, 151), {{100, 200, 0}, {111, 160, 0}, {150, 200, 0}}));
        CHECK(lapper.count(145, 151) == 3);
    }
    {  // test_seek_over_len  :1343-1353
        Lapper<uint32_t> lapper(nonoverlapping()), single({{10, 35, 0}});
        size_t cursor = 0;
        size_t hits = 0;
        for (const Iv &iv : lapper) hits += single.seek(iv.start, iv.stop, &cursor).size();
        CHECK(hits == 1);
    }
    {  // test_find_over_behind_first_match  :1357-1363, test_bad_skips  :1368-1385
        Lapper<uint32_t> lapper(badlapper());
        CHECK(*lapper.find(50, 55).front() == (Iv{50, 55, 0}));
        CHECK(lapper.find(50, 55).size() == lapper.count(50, 55));
        Lapper<uint32_t> bad({{25264912, 25264986, 0}, {27273024, 27273065, 0}, {27440273, 27440318, 0},
                              {27488033, 27488125, 0}, {27938410, 27938470, 0}, {27959118, 27959171, 0},
                              {28866309, 33141404, 0}});
        CHECK(same(bad.find(28974798, 33141355), {{28866309, 33141404, 0}}));
        CHECK(bad.count(28974798, 33141355) == 1);
    }
    {  // doctests: count/find :603-627, seek :646-655, cov :303-309, examples/ex1.rs:104-121
        Lapper<uint32_t> lapper(step_by(0, 100, 5, 2));
        CHECK(lapper.count(5, 11) == 2 && lapper.find(5, 11).size() == 2);
        size_t cursor = 0;
        for (const Iv &iv : lapper) CHECK(lapper.seek(iv.start, iv.stop, &cursor).size() == 1);
        CHECK(Lapper<uint32_t>(step_by(0, 20, 5, 10)).cov() == 25);
        Lapper<uint32_t> ex(badlapper()), other({{5, 15, 0}, {48, 80, 0}});
        ex.merge_overlaps();
        CHECK(ex.union_and_intersect(other) == std::make_pair((I)88, (I)27));
    }
    {  // test_depth_sanity / test_depth_hard  :1279-1313
        Lapper<uint32_t> l1({{0, 10, 0}, {5, 10, 0}});
        auto d = l1.depth();
        CHECK(d.size() == 2 && d[0] == (Iv{0, 5, 1}) && d[0].val == 1 && d[1] == (Iv{5, 10, 2}) && d[1].val == 2);
        Lapper<uint32_t> l2({{1, 10, 0}, {2, 5, 0}, {3, 8, 0}, {3, 8, 0}, {3, 8, 0}, {5, 8, 0}, {9, 11, 0}});
        const I want[6][3] = {{1, 2, 1}, {2, 3, 2}, {3, 8, 5}, {8, 9, 1}, {9, 10, 2}, {10, 11, 1}};
        auto d2 = l2.depth();
        CHECK(d2.size() == 6);
        for (size_t i = 0; i < 6 && i < d2.size(); i++)
            CHECK(d2[i].start == want[i][0] && d2[i].stop == want[i][1] && d2[i].val == want[i][2]);
    }
    {  // insert doctest :243-259 and insert_equality_overlapping :929-948
        Lapper<uint32_t> lapper({{0, 5, 1}, {6, 10, 2}});
        lapper.insert(Iv{0, 20, 5});
        CHECK(lapper.len() == 3);
        CHECK(same(lapper.find(1, 3), {{0, 5, 1}, {0, 20, 5}}));
        CHECK(lapper.find(1, 3)[1]->val == 5);
        Lapper<uint32_t> a(overlapping()), b(std::vector<Iv>{});
        for (const Iv &iv : overlapping()) b.insert(iv);
        CHECK(a.intervals.size() == b.intervals.size() && a.max_len() == b.max_len());
        for (size_t i = 0; i < a.intervals.size() && i < b.intervals.size(); i++) CHECK(a.intervals[i] == b.intervals[i]);
    }
    {  // serde_test  :1387-1409 (the `with_serde` feature): bincode round trip, then the same find / count
        Lapper<uint32_t> lapper({{25264912, 25264986, 0}, {27273024, 27273065, 0}, {27440273, 27440318, 0}, {27488033, 27488125, 0},
                                 {27938410, 27938470, 0}, {27959118, 27959171, 0}, {28866309, 33141404, 0}});
        const std::vector<uint8_t> serialized = lapper.to_bincode();
        Lapper<uint32_t> deserialized = Lapper<uint32_t>::from_bincode(serialized);
        CHECK(same(deserialized.find(28974798, 33141355), {{28866309, 33141404, 0}}));
        CHECK(deserialized.count(28974798, 33141355) == 1);
        CHECK(deserialized.to_bincode() == serialized);
        // layout, field by field (bincode 1.3 defaults): Lapper<I, u32> { intervals: [(5, 9, 7), (6, 20, 8)], starts: [5, 6],
        // stops: [9, 20], max_len: 14, cov: Some(15), overlaps_merged: false }
        Lapper<uint32_t> two({{6, 20, 8}, {5, 9, 7}});
        CHECK(two.set_cov() == 15);
        std::vector<uint8_t> want;
        auto le = [&want](uint64_t x, size_t w) { for (size_t i = 0; i < w; i++) want.push_back((uint8_t)(x >> (8 * i))); };
        const size_t W = sizeof(I);
        le(2, 8), le(5, W), le(9, W), le(7, 4), le(6, W), le(20, W), le(8, 4);
        le(2, 8), le(5, W), le(6, W), le(2, 8), le(9, W), le(20, W), le(14, W), le(1, 1), le(15, W), le(0, 1);
        CHECK(two.to_bincode() == want);
        // the cached fields survive: cov: Option<I> and overlaps_merged (src/lib.rs:119-122)
        Lapper<uint32_t> bad(badlapper());
        bad.set_cov();
        bad.merge_overlaps();
        Lapper<uint32_t> again = Lapper<uint32_t>::from_bincode(bad.to_bincode());
        CHECK(again.overlaps_merged() && again.cov() == bad.cov() && again.len() == 5);
        for (size_t i = 0; i < again.len() && i < bad.len(); i++) CHECK(again.intervals[i] == bad.intervals[i] && again.intervals[i].val == bad.intervals[i].val);
        // malformed input is an error, not a crash: truncated, trailing byte, bad Option tag, bad bool, absurd length
        std::vector<std::vector<uint8_t>> broken;
        broken.push_back(std::vector<uint8_t>(want.begin(), want.end() - 1));
        broken.push_back(want), broken.back().push_back(0);
        broken.push_back(want), broken.back()[want.size() - 2 - W] = 2;
        broken.push_back(want), broken.back().back() = 7;
        broken.push_back(std::vector<uint8_t>(8, 0xFF));
        broken.push_back(want), broken.back()[8] ^= 1;  // intervals[0].start no longer matches starts[0]
        for (const auto &b : broken) {
            bool threw = false;
            try {
                Lapper<uint32_t>::from_bincode(b);
            } catch (const lapper_b200::Error &e) {
                threw = e.code == LAPPER_ERR_INVALID_ARG;
            }
            CHECK(threw);
        }
    }
}

// Coordinate types narrower than 32 bits (`I: PrimInt + Unsigned`, src/lib.rs:88-100): the reference's small vectors
// through Lapper<u32, u8> / Lapper<u32, u16>, which widen into the u32 engine.
template <typename I>
static void small_coordinate_types() {
    typedef Interval<uint32_t, I> Iv;
    typedef lapper_b200::Lapper<uint32_t, I> L;
    const I top = std::numeric_limits<I>::max();
    auto step_by = [](I lo, I hi, I step, I len) {
        std::vector<Iv> v;
        for (unsigned x = lo; x < hi; x += step) v.push_back(Iv{(I)x, (I)(x + len), 0});
        return v;
    };
    L non(step_by(0, 100, 20, 10)), over(step_by(0, 100, 10, 15));
    size_t cursor = 0;
    CHECK(non.find(15, 20).empty() && non.count(15, 20) == 0 && non.seek(15, 20, &cursor).empty());   // :1023-1039
    CHECK(non.find(15, 25).size() == 1 && non.find(15, 25)[0]->start == 20 && non.count(15, 25) == 1);  // :1043-1056
    CHECK(over.find(8, 20).size() == 2 && over.count(8, 20) == 2 && over.find(8, 20)[1]->start == 10);  // :1102-1121
    CHECK(non.cov() == 50 && non.max_len() == 10);
    L bad({{70, 120, 0}, {10, 15, 1}, {10, 15, 2}, {12, 15, 3}, {14, 16, 4}, {40, 45, 5}, {50, 55, 6}, {60, 65, 7}, {68, 71, 8}, {70, 75, 9}});
    CHECK(bad.cov() == 73);
    L other({{5, 15, 0}, {48, 80, 0}});
    bad.merge_overlaps();                                                                           // :1124-1138
    CHECK(bad.len() == 5 && bad.intervals[0] == (Iv{10, 16, 0}) && bad.intervals[4] == (Iv{68, 120, 0}) && bad.intervals[0].val == 1);
    CHECK(bad.union_and_intersect(other) == std::make_pair((I)88, (I)27));                         // examples/ex1.rs:107-121
    auto d = L({{0, 10, 0}, {5, 10, 0}}).depth();                                                   // :1279-1290
    CHECK(d.size() == 2 && d[0].start == 0 && d[0].stop == 5 && d[0].val == 1 && d[1].start == 5 && d[1].stop == 10 && d[1].val == 2);
    // the top of the type: `start + 1` wraps to 0 for start = I::MAX (src/lib.rs:614, release builds), exactly as it
    // does in a Lapper<u32, _> for u32::MAX -- same set at the top of u32, same answers
    L hi({{(I)(top - 9), top, 0}, {(I)(top - 4), (I)(top - 1), 0}, {3, top, 0}});
    lapper_b200::Lapper<uint32_t, uint32_t> hi32({{0xFFFFFFFFu - 9, 0xFFFFFFFFu, 0}, {0xFFFFFFFFu - 4, 0xFFFFFFFFu - 1, 0}, {3, 0xFFFFFFFFu, 0}});
    for (unsigned a = 0; a <= 9; a++)
        for (unsigned b = 0; b <= 9; b++) {
            const I qa = (I)(top - a), qb = (I)(top - b);
            CHECK(hi.count(qa, qb) == hi32.count(0xFFFFFFFFu - a, 0xFFFFFFFFu - b));
            CHECK(hi.find(qa, qb).size() == hi32.find(0xFFFFFFFFu - a, 0xFFFFFFFFu - b).size());
        }
    // with_serde: the fields are as wide as I
    const std::vector<uint8_t> raw = over.to_bincode();
    CHECK(raw.size() == 8 + 10 * (2 * sizeof(I) + 4) + 2 * (8 + 10 * sizeof(I)) + sizeof(I) + 1 + 1);
    L back = L::from_bincode(raw);
    CHECK(back.to_bincode() == raw && back.count(8, 20) == 2);
    // insert keeps the reference's positions
    L ins({{0, 5, 1}, {6, 10, 2}});
    ins.insert(Iv{0, 20, 5});
    CHECK(ins.len() == 3 && ins.find(1, 3).size() == 2 && ins.find(1, 3)[1]->val == 5);
}

int main() {
    Suite<uint32_t>::run();
    const int f32 = failures;
    Suite<uint64_t>::run();
    const int f64 = failures;
    small_coordinate_types<uint16_t>();
    small_coordinate_types<uint8_t>();
    std::printf("I = u32: %d failures, I = u64: %d failures, I = u16 / u8: %d failures\n", f32, f64 - f32, failures - f64);
    std::printf(failures ? "%d FAILED\n" : "all reference tests passed (C++ facade)\n", failures);
    return failures ? 1 : 0;
}
