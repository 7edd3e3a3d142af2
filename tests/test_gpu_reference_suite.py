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
