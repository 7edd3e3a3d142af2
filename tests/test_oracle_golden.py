"""Pins the CPU oracle (oracle/lapper_oracle.c) to the reference before anything trusts it:
every known-answer test of /root/reference/src/lib.rs `mod tests`, the doctests and
examples/ex1.rs (restated in reference_suite.py), the insert_equality_* structural tests, the
Appendix-B edge cases against a second pure-Python transliteration, and a randomised differential
against the ten-line set definition."""
import numpy as np
import pytest

import reference_suite as rs


@pytest.fixture(scope="module")
def make(oracle):
    return lambda ivs: oracle.OracleLapper(ivs)


@pytest.mark.parametrize("check", rs.CORE_CHECKS + [rs.check_depth], ids=lambda f: f.__name__)
def test_reference_known_answers(make, check):
    check(make)


def test_interval_intersects(oracle):  # src/lib.rs:1181-1200
    i1, i2, i3, i4, i5 = (70, 120), (10, 15), (10, 15), (12, 15), (14, 16)
    i6, i7, i8, i9, i10 = (40, 50), (50, 55), (60, 65), (68, 71), (70, 75)
    f = oracle.interval_intersect
    assert f(i2, i3) == 5
    assert f(i2, i4) == 3
    assert f(i2, i5) == 1
    assert f(i9, i10) == 1
    assert f(i7, i8) == 0
    assert f(i6, i7) == 0
    assert f(i1, i10) == 5


@pytest.mark.parametrize(
    "data",
    [rs.nonoverlapping(), rs.overlapping(), rs.badlapper(), rs.single()],
    ids=["nonoverlapping", "overlapping", "badlapper", "single"],
)
def test_insert_equality(oracle, data):  # src/lib.rs:907-948, :976-1019
    new_lapper = oracle.OracleLapper(data)
    insert_lapper = oracle.OracleLapper([])
    for s, e in data:
        insert_lapper.insert(s, e)
    assert new_lapper.starts.tolist() == insert_lapper.starts.tolist()
    assert new_lapper.stops.tolist() == insert_lapper.stops.tolist()
    assert new_lapper.intervals == insert_lapper.intervals
    assert new_lapper.max_len == insert_lapper.max_len


def test_insert_equality_half_and_half(oracle):  # src/lib.rs:952-974
    data = [(x, x + 15) for x in range(100)]
    new_lapper = oracle.OracleLapper(data)
    insert_lapper = oracle.OracleLapper(data[:50])
    for s, e in reversed(data[50:]):
        insert_lapper.insert(s, e)
    assert new_lapper.starts.tolist() == insert_lapper.starts.tolist()
    assert new_lapper.stops.tolist() == insert_lapper.stops.tolist()
    assert new_lapper.intervals == insert_lapper.intervals
    assert new_lapper.max_len == insert_lapper.max_len


def test_insert_doctest(oracle):  # src/lib.rs:243-259
    lapper = oracle.OracleLapper([(0, 5), (6, 10)])
    lapper.insert(0, 20, 5)
    assert len(lapper) == 3
    assert lapper.find(1, 3) == [(0, 5), (0, 20)]


def test_edge_cases_match_transliteration(oracle):
    for ivs, (a, b) in rs.edge_case_vectors():
        lap = oracle.OracleLapper(ivs)
        assert lap.count(a, b) == rs.py_count(ivs, a, b), (ivs, a, b)
        assert lap.find_idx(a, b) == rs.py_find(ivs, a, b), (ivs, a, b)


def test_stable_ties_keep_input_order(oracle):
    # Vec::sort is stable (src/lib.rs:201): equal (start, stop) keep input order, visible through `val`
    ivs = [(5, 9), (1, 3), (5, 9), (5, 8), (1, 3), (5, 9)]
    lap = oracle.OracleLapper(ivs)
    s, e, v = lap.intervals_soa()
    assert list(zip(s.tolist(), e.tolist())) == sorted(ivs)
    assert v.tolist() == [1, 4, 3, 0, 2, 5]


@pytest.mark.parametrize("seed", range(8))
def test_differential_vs_set_definition(oracle, seed):
    rng = np.random.default_rng(seed)
    for _ in range(40):
        n = int(rng.integers(0, 60))
        universe = int(rng.choice([20, 200, 5000]))
        max_len = int(rng.choice([0, 3, 50, universe]))
        s, e, qs, qe = rs.random_case(rng, n, universe, max_len, 25, int(rng.choice([0, 1, 10, universe])))
        lap = oracle.OracleLapper(starts=s, stops=e)
        offsets, indices = lap.find_batch(qs, qe)
        counts = lap.count_batch(qs, qe)
        for j in range(qs.size):
            a, b = int(qs[j]), int(qe[j])
            want = oracle.brute_find(s, e, a, b)
            got = indices[int(offsets[j]):int(offsets[j + 1])].tolist()
            assert got == want
            assert lap.find_idx(a, b) == want
            assert int(counts[j]) == rs.py_count(list(zip(s.tolist(), e.tolist())), a, b)
            if a < b:  # Appendix A.4: |find| == count for non-empty queries over s<=e intervals
                assert int(counts[j]) == len(want)


@pytest.mark.parametrize("seed", range(4))
def test_seek_equals_find_on_sorted_queries(oracle, seed):  # src/lib.rs:641-645 precondition
    rng = np.random.default_rng(100 + seed)
    s, e, qs, qe = rs.random_case(rng, 500, 100000, 3000, 2000, 400)
    order = np.argsort(qs, kind="stable")
    qs, qe = qs[order], qe[order]
    lap = oracle.OracleLapper(starts=s, stops=e)
    offsets, indices = lap.find_batch(qs, qe)
    assert np.array_equal(lap.seek_batch(qs, qe, offsets), indices)


@pytest.mark.parametrize("seed", range(6))
def test_merge_cov_union_closed_forms(oracle, seed):
    """SURVEY Appendix A.5-A.7 closed forms vs the literal loops (these are what the GPU implements)."""
    rng = np.random.default_rng(200 + seed)
    for _ in range(30):
        n = int(rng.integers(1, 80))
        s, e, _, _ = rs.random_case(rng, n, 300, int(rng.choice([0, 5, 40])), 1, 1)
        lap = oracle.OracleLapper(starts=s, stops=e)
        ss, ee, vv = lap.intervals_soa()
        pmax = np.maximum.accumulate(ee.astype(np.int64))
        head = np.ones(n, bool)
        head[1:] = pmax[:-1] < ss[1:].astype(np.int64)
        last = np.ones(n, bool)
        last[:-1] = head[1:]
        merged = list(zip(ss[head].tolist(), pmax[last].tolist()))
        cov_closed = int(sum(b - a for a, b in merged))
        assert lap.cov() == cov_closed
        lap.merge_overlaps()
        assert lap.intervals == merged
        assert lap.intervals_soa()[2].tolist() == vv[head].tolist()  # val of a run = val of its head
        assert lap.cov() == cov_closed
        assert lap.starts.tolist() == [a for a, _ in merged] and lap.stops.tolist() == [b for _, b in merged]
        assert lap.max_len == max(b - a for a, b in merged)

        m = int(rng.integers(1, 60))
        s2, e2, _, _ = rs.random_case(rng, m, 300, int(rng.choice([0, 5, 40])), 1, 1)
        A, B = oracle.OracleLapper(starts=s, stops=e), oracle.OracleLapper(starts=s2, stops=e2)
        mark_a, mark_b = np.zeros(400, bool), np.zeros(400, bool)
        for a, b in zip(s.tolist(), e.tolist()):
            mark_a[a:b] = True
        for a, b in zip(s2.tolist(), e2.tolist()):
            mark_b[a:b] = True
        x = int((mark_a & mark_b).sum())
        want = (int(mark_a.sum()) + int(mark_b.sum()) - x, x)
        assert A.union_and_intersect(B) == want
        assert B.union_and_intersect(A) == want
        A.merge_overlaps()
        B.merge_overlaps()
        assert A.union_and_intersect(B) == want  # both-merged branch (src/lib.rs:535-543)
        B2 = oracle.OracleLapper(starts=s2, stops=e2)
        assert A.union_and_intersect(B2) == want  # mixed flags -> unmerged branch


def test_threads_agree(oracle):
    rng = np.random.default_rng(7)
    s, e, qs, qe = rs.random_case(rng, 5000, 1_000_000, 4000, 20000, 300)
    lap = oracle.OracleLapper(starts=s, stops=e)
    c1, c4 = lap.count_batch(qs, qe, 1), lap.count_batch(qs, qe, 4)
    assert np.array_equal(c1, c4)
    o1, i1 = lap.find_batch(qs, qe, 1)
    o4, i4 = lap.find_batch(qs, qe, 4)
    assert np.array_equal(o1, o4) and np.array_equal(i1, i4)
    assert lap.find_count_sum(qs, qe, 3) == int(o1[-1])


def depth_closed_form(s, e):
    """Closed form of IterDepth (src/lib.rs:735-784) that the GPU implements (csrc/setops.cu compute_depth)."""
    S, E = np.sort(np.asarray(s, np.int64)), np.sort(np.asarray(e, np.int64))
    U = np.unique(np.concatenate([S, E]))
    D = np.searchsorted(S, U, "right") - np.searchsorted(E, U, "right")
    has_start = np.searchsorted(S, U, "left") < np.searchsorted(S, U, "right")
    prev = np.concatenate([[-1], D[:-1]])
    change = prev != D
    zero = (D == 0) & (prev == 0) & has_start
    zero[0] = False
    ep = np.nonzero(change | zero)[0]
    out = []
    for q, t in enumerate(ep):
        if zero[t]:
            out.append((int(U[t]), int(U[t]), 0))
        elif D[t] > 0:
            out.append((int(U[t]), int(U[ep[q + 1]]), int(D[t])))
    return out


@pytest.mark.parametrize("seed", range(4))
def test_depth_closed_form(oracle, seed):
    rng = np.random.default_rng(400 + seed)
    for _ in range(300):
        n = int(rng.integers(1, 25))
        U = int(rng.choice([12, 40, 200]))
        s = rng.integers(0, U, n)
        ln = rng.integers(0, int(rng.choice([1, 2, 6, 30])), n)
        if rng.random() < 0.5:
            ln[rng.random(n) < 0.4] = 0  # zero-length intervals: the IterDepth quirks live here
        assert oracle.OracleLapper(starts=s, stops=s + ln).depth() == depth_closed_form(s, s + ln)
