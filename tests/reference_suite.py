"""The reference's own known-answer tests, restated once and run against any Lapper-like class.

Each function cites the reference test it transcribes (/root/reference/src/lib.rs `mod tests`,
the doctests and examples/ex1.rs).  `make(intervals)` builds a lapper from (start, stop) tuples;
the object must offer the reference's method names: find / seek / count / merge_overlaps / cov /
set_cov / union_and_intersect / intervals / starts / stops / max_len (+ depth, insert where the
implementation has them).  find/seek return lists of (start, stop) in iterator order; the seek
cursor (`&mut usize`) is a one-element list.

Run against the CPU oracle in test_oracle_golden.py (pins the oracle) and against the CUDA
engine through the C ABI in test_gpu_reference_suite.py (pins the product).
"""
from __future__ import annotations

import numpy as np


def nonoverlapping():  # src/lib.rs:857-867
    return [(x, x + 10) for x in range(0, 100, 20)]


def overlapping():  # src/lib.rs:869-879
    return [(x, x + 15) for x in range(0, 100, 10)]


def badlapper():  # src/lib.rs:881-895
    return [(70, 120), (10, 15), (10, 15), (12, 15), (14, 16), (40, 45), (50, 55), (60, 65), (68, 71), (70, 75)]


def single():  # src/lib.rs:897-904
    return [(10, 35)]


def _tl(a):
    return [int(x) for x in a]


# ---- boundary semantics: src/lib.rs:1023-1099 -------------------------------------------------
def check_query_stop_interval_start(make):  # :1023-1029
    lapper = make(nonoverlapping())
    cursor = [0]
    assert lapper.find(15, 20) == []
    assert lapper.seek(15, 20, cursor) == []
    assert len(lapper.find(15, 20)) == lapper.count(15, 20)


def check_query_start_interval_stop(make):  # :1033-1039
    lapper = make(nonoverlapping())
    cursor = [0]
    assert lapper.find(30, 35) == []
    assert lapper.seek(30, 35, cursor) == []
    assert len(lapper.find(30, 35)) == lapper.count(30, 35)


def _expect_20_30(make, a, b):
    lapper = make(nonoverlapping())
    cursor = [0]
    expected = (20, 30)
    assert lapper.find(a, b)[0] == expected
    assert lapper.seek(a, b, cursor)[0] == expected
    assert len(lapper.find(a, b)) == lapper.count(a, b)


def check_query_overlaps_interval_start(make):  # :1043-1054
    _expect_20_30(make, 15, 25)


def check_query_overlaps_interval_stop(make):  # :1058-1069
    _expect_20_30(make, 25, 35)


def check_interval_envelops_query(make):  # :1073-1084
    _expect_20_30(make, 22, 27)


def check_query_envolops_interval(make):  # :1088-1099
    _expect_20_30(make, 15, 35)


# ---- order of results -------------------------------------------------------------------------
def check_overlapping_intervals(make):  # :1102-1121
    lapper = make(overlapping())
    cursor = [0]
    e1, e2 = (0, 15), (10, 25)
    assert lapper.find(8, 20) == [e1, e2]
    assert lapper.seek(8, 20, cursor) == [e1, e2]
    assert lapper.count(8, 20) == 2


def check_find_overlaps_in_large_intervals(make):  # :1243-1276
    data = [(0, 8), (1, 10), (2, 5), (3, 8), (4, 7), (5, 8), (8, 8), (9, 11), (10, 13), (100, 200),
            (110, 120), (110, 124), (111, 160), (150, 200)]
    lapper = make(data)
    assert lapper.find(8, 11) == [(1, 10), (9, 11), (10, 13)]
    assert lapper.count(8, 11) == 3
    assert lapper.find(145, 151) == [(100, 200), (111, 160), (150, 200)]
    assert lapper.count(145, 151) == 3


# ---- merge / cov / set algebra ----------------------------------------------------------------
def check_merge_overlaps(make):  # :1124-1138
    lapper = make(badlapper())
    expected = [(10, 16), (40, 45), (50, 55), (60, 65), (68, 120)]
    assert len(lapper.intervals) == len(lapper.starts)
    lapper.merge_overlaps()
    assert lapper.intervals == expected
    assert len(lapper.intervals) == len(lapper.starts)
    assert lapper.overlaps_merged


def check_merge_overlaps_find(make):  # :1143-1165
    lapper = make([(2, 3), (3, 4), (4, 6), (6, 7), (7, 8)])
    assert lapper.find(7, 9) == [(7, 8)]
    lapper.merge_overlaps()
    assert lapper.find(7, 9) == [(2, 8)]


def check_lapper_cov(make):  # :1168-1178 (+ examples/ex1.rs:104 pins the value 73)
    lapper = make(badlapper())
    before = lapper.cov()
    lapper.merge_overlaps()
    after = lapper.cov()
    assert before == after == 73
    lapper = make(nonoverlapping())
    lapper.set_cov()
    assert lapper.cov() == 50


def check_union_and_intersect(make):  # :1203-1240 and doctest :472-511
    data1 = [(70, 120), (10, 15), (12, 15), (14, 16), (68, 71)]
    data2 = [(10, 15), (40, 45), (50, 55), (60, 65), (70, 75)]
    lapper1, lapper2 = make(data1), make(data2)
    assert lapper1.union_and_intersect(lapper2) == (73, 10)
    assert lapper2.union_and_intersect(lapper1) == (73, 10)
    lapper1.merge_overlaps()
    lapper1.set_cov()
    lapper2.merge_overlaps()
    lapper2.set_cov()
    assert lapper1.union_and_intersect(lapper2) == (73, 10)
    assert lapper2.union_and_intersect(lapper1) == (73, 10)


def check_example_ex1(make):  # examples/ex1.rs:65-121, README.md:160-217
    lapper = make(badlapper())
    assert lapper.find(11, 15) == [(10, 15), (10, 15), (12, 15), (14, 16)]
    lapper.merge_overlaps()
    assert lapper.find(11, 15) == [(10, 16)]
    assert lapper.cov() == 73
    other = make([(5, 15), (48, 80)])
    assert lapper.union_and_intersect(other) == (88, 27)


# ---- regression tests "from real life" --------------------------------------------------------
def check_seek_over_len(make):  # :1343-1353
    lapper = make(nonoverlapping())
    one = make(single())
    cursor = [0]
    got = []
    for (s, e) in lapper.intervals:
        got.append(one.seek(s, e, cursor))
    # (10,35) overlaps [0,10)? no; [20,30) yes; [40,50).. no.  The point of the reference test is "no panic".
    assert got == [[], [(10, 35)], [], [], []]


def check_find_over_behind_first_match(make):  # :1357-1363
    lapper = make(badlapper())
    assert lapper.find(50, 55)[0] == (50, 55)
    assert len(lapper.find(50, 55)) == lapper.count(50, 55)


def check_bad_skips(make):  # :1368-1385
    data = [(25264912, 25264986), (27273024, 27273065), (27440273, 27440318), (27488033, 27488125),
            (27938410, 27938470), (27959118, 27959171), (28866309, 33141404)]
    lapper = make(data)
    assert lapper.find(28974798, 33141355) == [(28866309, 33141404)]
    assert lapper.count(28974798, 33141355) == 1


# ---- doctests ---------------------------------------------------------------------------------
def check_doctest_crate(make):  # src/lib.rs:51-77
    laps = make([(x, x + 2) for x in range(0, 20, 5)])
    assert laps.find(6, 11)[0] == (5, 7)
    sim = 0
    cursor = [0]
    for i in range(0, 10, 3):
        sim += sum(min(i + 2, e) - max(i, s) for (s, e) in laps.seek(i, i + 2, cursor))
    assert sim == 4


def check_doctest_cov(make):  # :303-309
    lapper = make([(x, x + 10) for x in range(0, 20, 5)])
    assert len(lapper.intervals) == 4  # :276-283
    assert lapper.cov() == 25


def check_doctest_count_find_seek(make):  # :603-609, :621-627, :646-655
    lapper = make([(x, x + 2) for x in range(0, 100, 5)])
    assert lapper.count(5, 11) == 2
    assert len(lapper.find(5, 11)) == 2
    cursor = [0]
    for (s, e) in lapper.intervals:
        assert len(lapper.seek(s, e, cursor)) == 1


def check_doctest_empty(make):  # :290-295 + SURVEY Appendix B "empty lapper"
    lapper = make([])
    assert len(lapper.intervals) == 0
    assert lapper.count(0, 10) == 0
    assert lapper.find(0, 10) == []
    assert lapper.cov() == 0
    lapper.merge_overlaps()
    assert not lapper.overlaps_merged  # :366, :382 -- the flag is only set when there is a first interval
    assert lapper.find(0, 10) == []


# ---- build structure: what insert_equality_* compare (src/lib.rs:907-1019) ---------------------
def check_build_structure(make):
    for data in (nonoverlapping(), overlapping(), [(x, x + 15) for x in range(100)], badlapper(), single()):
        lapper = make(data)
        s = sorted(data)
        assert lapper.intervals == s
        assert _tl(lapper.starts) == sorted(a for a, _ in data)
        assert _tl(lapper.stops) == sorted(b for _, b in data)
        assert lapper.max_len == max(b - a for a, b in data)


# ---- depth (src/lib.rs:1279-1337, doctest :565-575) -- only for implementations that have it ---
def check_depth(make):
    assert make([(0, 10), (5, 10)]).depth() == [(0, 5, 1), (5, 10, 2)]
    hard = [(1, 10), (2, 5), (3, 8), (3, 8), (3, 8), (5, 8), (9, 11)]
    exp = [(1, 2, 1), (2, 3, 2), (3, 8, 5), (8, 9, 1), (9, 10, 2), (10, 11, 1)]
    assert make(hard).depth() == exp
    assert make(hard + [(15, 20)]).depth() == exp + [(15, 20, 1)]
    assert make([(x, x + 10) for x in range(0, 20, 5)]).depth() == [(0, 5, 1), (5, 20, 2), (20, 25, 1)]


CORE_CHECKS = [
    check_query_stop_interval_start,
    check_query_start_interval_stop,
    check_query_overlaps_interval_start,
    check_query_overlaps_interval_stop,
    check_interval_envelops_query,
    check_query_envolops_interval,
    check_overlapping_intervals,
    check_find_overlaps_in_large_intervals,
    check_merge_overlaps,
    check_merge_overlaps_find,
    check_lapper_cov,
    check_union_and_intersect,
    check_example_ex1,
    check_seek_over_len,
    check_find_over_behind_first_match,
    check_bad_skips,
    check_doctest_crate,
    check_doctest_cov,
    check_doctest_count_find_seek,
    check_doctest_empty,
    check_build_structure,
]


# ---- Appendix B edge cases, stated as the reference's release-mode behaviour --------------------
def py_bsearch_seq(key, elems):
    """Pure-Python transliteration of bsearch_seq_ref (src/lib.rs:448-466); second opinion for the oracle."""
    n = len(elems)
    if n == 0 or elems[0] >= key:
        return 0
    if elems[n - 1] < key:
        return n
    cursor, length = 0, n
    while length > 1:
        half = length >> 1
        length -= half
        cursor += int(elems[cursor + half - 1] < key) * half
    return cursor


def py_count(intervals, a, b, bits=32):
    """Lapper::count (src/lib.rs:611-618) with `start + 1` wrapping in I (u32; bits=64: u64) and usize subtraction
    wrapping (release build)."""
    starts = sorted(s for s, _ in intervals)
    stops = sorted(e for _, e in intervals)
    n = len(intervals)
    first = py_bsearch_seq((a + 1) & ((1 << bits) - 1), stops)
    last = py_bsearch_seq(b, starts)
    return (n - first - (n - last)) % (1 << 64)


def py_find(intervals, a, b):
    """find + IterFind::next (src/lib.rs:629-639, :704-717), literal, returning sorted positions."""
    ivs = sorted(intervals)  # (start, stop); Python's sort is stable like Vec::sort
    max_len = max([e - s if e >= s else 0 for s, e in ivs], default=0)
    key = a - max_len if a >= max_len else 0
    size, low = len(ivs), 0
    while size > 0:  # lower_bound :423-437
        half = size // 2
        other_half = size - half
        probe, other_low = low + half, low + other_half
        size = half
        low = other_low if ivs[probe][0] < key else low
    out, off = [], low
    while off < len(ivs):
        s, e = ivs[off]
        off += 1
        if s < b and e > a:
            out.append(off - 1)
        elif s >= b:
            break
    return out


def edge_case_vectors(bits=32):
    """(intervals, query) pairs from SURVEY Appendix B; expectations come from py_count / py_find.  X = I::MAX."""
    X = (1 << bits) - 1
    return [
        ([(8, 8)], (7, 9)),                 # zero-length interval strictly straddled
        ([(5, 10)], (7, 7)),                # zero-length query strictly inside
        ([(7, 7)], (7, 7)),                 # both zero-length: count wraps to usize::MAX (src/lib.rs:616-617)
        ([(10, 5)], (6, 8)),                # inverted interval: never found, count wraps
        ([(5, 10)], (9, 6)),                # inverted query: raw predicate
        ([(5, 10), (20, 30)], (25, 8)),
        ([(0, X)], (X, X)),                 # start+1 wraps to 0 (src/lib.rs:614)
        ([(0, X), (5, 6)], (X, 0)),
        ([(3, X)], (X - 1, X)),
        ([(0, 0)], (0, 0)),
        ([(0, 0), (0, 1), (1, 1)], (0, 1)),
        ([(X, X), (X - 1, X)], (X - 1, X)),
        ([(1, 4), (2, 3), (2, 3), (2, 3)], (0, X)),
        ([(10, 5), (3, 20), (7, 2)], (4, 6)),
    ]


def random_case(rng: np.random.Generator, n: int, universe: int, max_len: int, nq: int, qlen: int):
    s = rng.integers(0, universe, n, dtype=np.int64)
    ln = rng.integers(0, max_len + 1, n, dtype=np.int64)
    e = np.minimum(s + ln, 0xFFFFFFFF)
    qs = rng.integers(0, universe + max_len, nq, dtype=np.int64)
    qe = np.minimum(qs + rng.integers(0, qlen + 1, nq, dtype=np.int64), 0xFFFFFFFF)
    return s.astype(np.uint32), e.astype(np.uint32), qs.astype(np.uint32), qe.astype(np.uint32)
