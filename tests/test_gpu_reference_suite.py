"""The reference's own tests (reference_suite.py) run against the CUDA engine through the C ABI,
on both search paths (BITS bucket sectors / direct bisection under a shared-memory directory)."""
import pytest

import reference_suite as rs

pytestmark = pytest.mark.gpu

FLAG_SETS = {"auto": 0, "buckets": 2, "buckets_shift0": 2 | ((0 + 1) << 8), "buckets_shift3": (3 + 1) << 8, "plain": 1}


@pytest.fixture(scope="module", params=list(FLAG_SETS))
def make(request):
    from rust_lapper_b200 import Lapper

    flags = FLAG_SETS[request.param]
    return lambda ivs: Lapper(ivs, flags=flags)


@pytest.mark.parametrize("check", rs.CORE_CHECKS, ids=lambda f: f.__name__)
def test_reference_known_answers(make, check):
    check(make)


def test_edge_cases_match_reference_semantics(make):
    for ivs, (a, b) in rs.edge_case_vectors():
        lap = make(ivs)
        assert lap.count(a, b) == rs.py_count(ivs, a, b), (ivs, a, b)
        assert lap.find_idx(a, b) == rs.py_find(ivs, a, b), (ivs, a, b)


def test_stable_ties_keep_input_order(make):
    ivs = [(5, 9), (1, 3), (5, 9), (5, 8), (1, 3), (5, 9)]
    lap = make(ivs)
    s, e, p = lap.intervals_soa()
    assert list(zip(s.tolist(), e.tolist())) == sorted(ivs)
    assert p.tolist() == [1, 4, 3, 0, 2, 5]


@pytest.mark.parametrize(
    "data",
    [rs.nonoverlapping(), rs.overlapping(), rs.badlapper(), rs.single()],
    ids=["nonoverlapping", "overlapping", "badlapper", "single"],
)
def test_insert_equality(make, data):  # src/lib.rs:907-948, :976-1019
    new_lapper = make(data)
    insert_lapper = make([])
    for s, e in data:
        insert_lapper.insert(s, e)
    assert new_lapper.starts.tolist() == insert_lapper.starts.tolist()
    assert new_lapper.stops.tolist() == insert_lapper.stops.tolist()
    assert new_lapper.intervals == insert_lapper.intervals
    assert new_lapper.max_len == insert_lapper.max_len


def test_insert_equality_half_and_half(make):  # src/lib.rs:952-974
    data = [(x, x + 15) for x in range(100)]
    new_lapper = make(data)
    insert_lapper = make(data[:50])
    for s, e in reversed(data[50:]):
        insert_lapper.insert(s, e)
    assert new_lapper.starts.tolist() == insert_lapper.starts.tolist()
    assert new_lapper.stops.tolist() == insert_lapper.stops.tolist()
    assert new_lapper.intervals == insert_lapper.intervals
    assert new_lapper.max_len == insert_lapper.max_len


def test_insert_doctest_and_side_effects(make, oracle):  # src/lib.rs:243-259, :271-272
    lapper = make([(0, 5), (6, 10)])
    lapper.merge_overlaps()
    lapper.set_cov()
    lapper.insert(0, 20)
    assert len(lapper.intervals) == 3
    assert lapper.find(1, 3) == [(0, 5), (0, 20)]
    assert not lapper.overlaps_merged
    assert lapper.cov() == 20
    # a duplicate goes BEFORE the equal intervals already there (bsearch_seq_ref), like the oracle
    cpu = oracle.OracleLapper([(5, 9), (5, 9), (1, 3)])
    gpu = make([(5, 9), (5, 9), (1, 3)])
    cpu.insert(5, 9, 3)
    gpu.insert(5, 9)
    assert gpu.intervals_soa()[2].tolist() == cpu.intervals_soa()[2].tolist() == [2, 3, 0, 1]


def test_depth_reference_vectors(make):  # src/lib.rs:1279-1337 and the doctest :565-575
    rs.check_depth(make)


def test_depth_matches_oracle_on_random_sets(make, oracle):
    import numpy as np

    rng = np.random.default_rng(77)
    for _ in range(25):
        n = int(rng.integers(1, 60))
        U = int(rng.choice([15, 60, 400]))
        s = rng.integers(0, U, n)
        ln = rng.integers(0, int(rng.choice([1, 3, 9, 50])), n)
        if rng.random() < 0.5:
            ln[rng.random(n) < 0.4] = 0
        ivs = list(zip(s.tolist(), (s + ln).tolist()))
        assert make(ivs).depth() == oracle.OracleLapper(ivs).depth()
    with pytest.raises(IndexError):
        make([]).depth()
