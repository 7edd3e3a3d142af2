"""The u64 engine (lapper64_* / rust_lapper_b200.Lapper64, csrc/engine64.cu) against the same checker as the u32
engine: the reference's known-answer tests (reference_suite.py), and the C oracle built with I = u64
(oracle/liblapper_oracle64.so) on seeded random sets whose coordinates do not fit 32 bits -- counts (u64, wrapping),
CSR offsets and positions, the sorted structure incl. tie order, cov / merge_overlaps / union_and_intersect, the
capacity protocol of the one-call device entry point -- always bit for bit."""
import numpy as np
import pytest

import reference_suite as rs

pytestmark = pytest.mark.gpu

U64_MAX = (1 << 64) - 1


@pytest.fixture(scope="module")
def oracle64():
    from oracle import lapper_oracle

    lapper_oracle.load(coord_bits=64)
    return lambda **kw: lapper_oracle.OracleLapper(coord_bits=64, **kw)


@pytest.fixture(scope="module")
def make():
    from rust_lapper_b200 import Lapper64

    return lambda ivs: Lapper64(ivs)


@pytest.mark.parametrize("check", rs.CORE_CHECKS + [rs.check_depth], ids=lambda f: f.__name__)
def test_reference_known_answers_u64(make, check):
    check(make)


@pytest.mark.parametrize(
    "data",
    [rs.nonoverlapping(), rs.overlapping(), rs.badlapper(), rs.single()],
    ids=["nonoverlapping", "overlapping", "badlapper", "single"],
)
def test_insert_equality_u64(make, data):  # src/lib.rs:907-948, :976-1019
    new_lapper = make(data)
    insert_lapper = make([])
    for s, e in data:
        insert_lapper.insert(s, e)
    assert new_lapper.starts.tolist() == insert_lapper.starts.tolist()
    assert new_lapper.stops.tolist() == insert_lapper.stops.tolist()
    assert new_lapper.intervals == insert_lapper.intervals
    assert new_lapper.max_len == insert_lapper.max_len


def test_insert_depth_seek_against_the_oracle_u64(oracle64):
    """insert (positions before equal intervals, perm of the new interval), depth and the seek cursor protocol on
    coordinates beyond 32 bits, against the oracle built with I = u64"""
    from rust_lapper_b200 import Lapper64

    rng = np.random.default_rng(17)
    sh = 1 << 41
    s = (rng.integers(0, 5000, 400).astype(np.uint64) + np.uint64(sh))
    e = s + rng.integers(0, 300, 400).astype(np.uint64)
    gpu, cpu = Lapper64(starts=s[:300], stops=e[:300]), oracle64(starts=s[:300], stops=e[:300])
    for a, b in zip(s[300:].tolist(), e[300:].tolist()):
        gpu.insert(a, b)
        cpu.insert(a, b, len(cpu))  # val = input position, like the engine's perm
    gs, ge, gp = gpu.intervals_soa()
    cs, ce, cv = cpu.intervals_soa()
    assert np.array_equal(gs, cs) and np.array_equal(ge, ce) and np.array_equal(gp, cv)
    assert np.array_equal(gpu.stops, cpu.stops) and gpu.max_len == cpu.max_len
    assert gpu.depth() == cpu.depth()
    qs = np.sort(rng.integers(sh - 100, sh + 5400, 300).astype(np.uint64))
    cg, cc = [0], [0]
    for a in qs.tolist():
        assert gpu.seek_idx(a, a + 40, cg) == cpu.seek_idx(a, a + 40, cc)
        assert cg == cc
    # unsorted queries: the cursor protocol differs from find and must still match the reference
    cg, cc = [0], [0]
    for a in rng.permutation(qs).tolist()[:150]:
        assert gpu.seek_idx(a, a + 40, cg) == cpu.seek_idx(a, a + 40, cc)
        assert cg == cc


def test_reference_known_answers_shifted_beyond_32_bits(make):
    """the same vectors translated by 2^40: every coordinate needs more than 32 bits"""
    sh = 1 << 40
    ivs = [(s + sh, e + sh) for s, e in rs.badlapper()]
    lap = make(ivs)
    assert lap.find(15 + sh, 20 + sh) == [(14 + sh, 16 + sh)]  # src/lib.rs:1368-1385 shape
    assert lap.count(8 + sh, 20 + sh) == 4
    assert lap.find(0, sh) == [] and lap.count(0, sh) == 0
    assert lap.cov() == 73  # examples/ex1.rs:104, translation-invariant
    lap.merge_overlaps()
    assert lap.intervals == [(10 + sh, 16 + sh), (40 + sh, 45 + sh), (50 + sh, 55 + sh), (60 + sh, 65 + sh), (68 + sh, 120 + sh)]


def test_edge_cases_match_reference_semantics_u64(make):
    for ivs, (a, b) in rs.edge_case_vectors(64) + rs.edge_case_vectors(32):
        lap = make(ivs)
        assert lap.count(a, b) == rs.py_count(ivs, a, b, 64), (ivs, a, b)
        assert lap.find_idx(a, b) == rs.py_find(ivs, a, b), (ivs, a, b)


def test_stable_ties_keep_input_order_u64(make):
    ivs = [(5, 9), (1, 3), (5, 9), (5, 8), (1, 3), (5, 9)]
    s, e, p = make(ivs).intervals_soa()
    assert list(zip(s.tolist(), e.tolist())) == sorted(ivs)
    assert p.tolist() == [1, 4, 3, 0, 2, 5]


def _random_set(rng, n, universe, max_len, long_frac=0.0):
    s = rng.integers(0, universe, n, dtype=np.uint64)
    ln = rng.integers(0, max_len, n, dtype=np.uint64)
    if long_frac:
        k = rng.random(n) < long_frac
        ln[k] = rng.integers(universe // 4, universe // 2, int(k.sum()), dtype=np.uint64)
    return s, s + ln


def _same(gpu, cpu, qs, qe):
    assert np.array_equal(gpu.count_batch(qs, qe), cpu.count_batch(qs, qe, 4))
    go, gi = gpu.find_batch(qs, qe)
    co, ci = cpu.find_batch(qs, qe, 4)
    assert np.array_equal(go, co)
    assert np.array_equal(gi, ci)


@pytest.mark.parametrize("n,universe,max_len,long_frac", [
    (1, 1 << 50, 1 << 20, 0.0),
    (1000, 1 << 45, 1 << 30, 0.0),
    (200_000, 1 << 60, 1 << 44, 0.0),
    (200_000, 1 << 40, 1 << 24, 0.01),   # a few very long intervals: max_len stretches the reference's window
    (50_000, 1 << 20, 1 << 8, 0.0),      # many duplicates and ties
    (2049, 3_000_000_000, 5000, 0.0),    # sizes around the scan tile
])
def test_random_sets_u64(oracle64, n, universe, max_len, long_frac):
    from rust_lapper_b200 import Lapper64

    rng = np.random.default_rng(n ^ 0xABCD)
    s, e = _random_set(rng, n, universe, max_len, long_frac)
    gpu, cpu = Lapper64(starts=s, stops=e), oracle64(starts=s, stops=e)
    gs, ge, gp = gpu.intervals_soa()
    cs, ce, cv = cpu.intervals_soa()
    assert np.array_equal(gs, cs) and np.array_equal(ge, ce) and np.array_equal(gp, cv)  # val = input position
    assert np.array_equal(gpu.stops, cpu.stops) and gpu.max_len == cpu.max_len
    nq = 30_000
    qs = rng.integers(0, universe + universe // 8, nq, dtype=np.uint64)
    qe = qs + rng.integers(0, max(max_len, 2), nq, dtype=np.uint64)
    # a tenth of the queries empty (a == b) or inverted, some at the ends of the coordinate range
    k = rng.choice(nq, nq // 10, replace=False)
    qe[k] = qs[k] - rng.integers(0, 3, k.size, dtype=np.uint64).clip(max=qs[k])
    qs[:4] = [0, 0, U64_MAX, U64_MAX - 5]
    qe[:4] = [0, U64_MAX, U64_MAX, U64_MAX]
    _same(gpu, cpu, qs, qe)
    _same(gpu, cpu, np.sort(qs), np.sort(qs) + np.uint64(150))
    assert gpu.cov() == cpu.cov()
    assert gpu.set_cov() == cpu.set_cov()


def test_coordinates_at_the_ends_of_u64(oracle64):
    from rust_lapper_b200 import Lapper64

    s = np.array([0, 0, 1, U64_MAX - 10, U64_MAX - 1, U64_MAX, 1 << 63, (1 << 63) - 1], dtype=np.uint64)
    e = np.array([0, U64_MAX, 1, U64_MAX, U64_MAX, U64_MAX, (1 << 63) + 5, 1 << 63], dtype=np.uint64)
    gpu, cpu = Lapper64(starts=s, stops=e), oracle64(starts=s, stops=e)
    pts = np.array([0, 1, 2, (1 << 63) - 1, 1 << 63, (1 << 63) + 5, U64_MAX - 10, U64_MAX - 1, U64_MAX], dtype=np.uint64)
    qs, qe = np.meshgrid(pts, pts)
    _same(gpu, cpu, qs.reshape(-1).copy(), qe.reshape(-1).copy())
    assert gpu.max_len == cpu.max_len == U64_MAX


def test_inverted_intervals_u64(oracle64):
    """count, find and merge_overlaps have no precondition on the intervals (src/lib.rs:361-416, :610-639); cov / union refuse"""
    from rust_lapper_b200 import Lapper64, LapperError, _ffi

    rng = np.random.default_rng(3)
    s, e = _random_set(rng, 20_000, 1 << 42, 1 << 28)
    k = rng.choice(s.size, 500, replace=False)
    s[k], e[k] = e[k], s[k]
    gpu, cpu = Lapper64(starts=s, stops=e), oracle64(starts=s, stops=e)
    assert gpu.has_inverted
    qs = rng.integers(0, 1 << 42, 20_000, dtype=np.uint64)
    _same(gpu, cpu, qs, qs + rng.integers(0, 1 << 29, qs.size, dtype=np.uint64))
    with pytest.raises(LapperError) as ei:
        gpu.cov()
    assert ei.value.code == _ffi.ERR_PRECONDITION
    # merge_overlaps follows the reference's loop (comparisons only) on such sets
    gpu.merge_overlaps(), cpu.merge_overlaps()
    gs, ge, gp = gpu.intervals_soa()
    cs, ce, cv = cpu.intervals_soa()
    assert np.array_equal(gs, cs) and np.array_equal(ge, ce) and np.array_equal(gp, cv)
    assert np.array_equal(gpu.stops, cpu.stops) and gpu.max_len == cpu.max_len
    _same(gpu, cpu, qs, qs + rng.integers(0, 1 << 29, qs.size, dtype=np.uint64))


def test_empty_u64():
    from rust_lapper_b200 import Lapper64

    lap = Lapper64([])
    assert len(lap) == 0 and lap.is_empty() and lap.max_len == 0
    assert lap.count(1, 100) == 0 and lap.find(1, 100) == []
    assert lap.count_batch([], []).size == 0
    off, idx = lap.find_batch([5, 6], [7, 8])
    assert off.tolist() == [0, 0, 0] and idx.size == 0
    assert lap.cov() == 0
    lap.merge_overlaps()
    assert not lap.overlaps_merged  # src/lib.rs:366: set only when there is a first interval


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_set_operations_u64(oracle64, seed):
    from rust_lapper_b200 import Lapper64

    rng = np.random.default_rng(seed)
    U = 1 << 44
    a_s, a_e = _random_set(rng, 30_000, U, 1 << 31, 0.001)
    b_s, b_e = _random_set(rng, 20_000, U, 1 << 30)
    ga, gb = Lapper64(starts=a_s, stops=a_e), Lapper64(starts=b_s, stops=b_e)
    ca, cb = oracle64(starts=a_s, stops=a_e), oracle64(starts=b_s, stops=b_e)
    assert ga.cov() == ca.cov() and gb.cov() == cb.cov()
    # the reference's unmerged branch materialises every pairwise intersection: keep the oracle's work bounded by
    # comparing on the merged sets (both branches give the same values, src/lib.rs:497-510)
    ga.merge_overlaps(), gb.merge_overlaps(), ca.merge_overlaps(), cb.merge_overlaps()
    for g, c in ((ga, ca), (gb, cb)):
        gs, ge, gp = g.intervals_soa()
        cs, ce, cv = c.intervals_soa()
        assert np.array_equal(gs, cs) and np.array_equal(ge, ce) and np.array_equal(gp, cv)
        assert np.array_equal(g.stops, c.stops) and g.max_len == c.max_len
        assert g.overlaps_merged and c.overlaps_merged
    assert ga.union_and_intersect(gb) == ca.union_and_intersect(cb)
    assert gb.union_and_intersect(ga) == cb.union_and_intersect(ca)
    ga.set_cov(), ca.set_cov()
    assert ga.union_and_intersect(gb) == ca.union_and_intersect(cb)
    qs = np.sort(rng.integers(0, U, 20_000, dtype=np.uint64))
    _same(ga, ca, qs, qs + np.uint64(1 << 20))


def test_find_one_call_device_api_u64(oracle64):
    """lapper64_find_batch_dev: exact fit, room to spare, too small (LAPPER_ERR_CAPACITY, offsets complete, positions
    untouched), the size query; lapper64_count_batch_dev on device buffers"""
    import ctypes as C

    import torch

    from rust_lapper_b200 import _ffi

    lib = _ffi.lib()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(9)
    s, e = _random_set(rng, 100_000, 1 << 46, 1 << 32)
    qs = np.sort(rng.integers(0, 1 << 46, 40_000, dtype=np.uint64))
    qe = qs + np.uint64(1 << 30)
    cpu = oracle64(starts=s, stops=e)
    want_o, want_i = cpu.find_batch(qs, qe, 4)
    total = int(want_o[-1])
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    ts, te = torch.from_numpy(s.view(np.int64)).to(dev), torch.from_numpy(e.view(np.int64)).to(dev)
    tq, tqe = torch.from_numpy(qs.view(np.int64)).to(dev), torch.from_numpy(qe.view(np.int64)).to(dev)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    h = C.c_void_p(0)
    _ffi.check(lib.lapper64_build_dev(vp(ts), vp(te), len(s), 0, sp, C.byref(h)))
    try:
        nq = len(qs)
        cnt = torch.empty(nq, dtype=torch.int64, device=dev)
        _ffi.check(lib.lapper64_count_batch_dev(h, vp(tq), vp(tqe), nq, vp(cnt), sp))
        torch.cuda.synchronize()
        assert np.array_equal(cnt.cpu().numpy().view(np.uint64), cpu.count_batch(qs, qe, 4))
        for cap, want_st in ((total, _ffi.OK), (total + 999, _ffi.OK), (total - 1, _ffi.ERR_CAPACITY)):
            off = torch.empty(nq + 1, dtype=torch.int64, device=dev)
            idx = torch.full((cap,), 0x5A5A5A5A, dtype=torch.int32, device=dev)
            t = C.c_uint64(0)
            st = lib.lapper64_find_batch_dev(h, vp(tq), vp(tqe), nq, vp(off), vp(idx), cap, C.byref(t), sp)
            torch.cuda.synchronize()
            assert st == want_st and int(t.value) == total
            assert np.array_equal(off.cpu().numpy().view(np.uint64), want_o)
            got = idx.cpu().numpy().view(np.uint32)
            if st == _ffi.OK:
                assert np.array_equal(got[:total], want_i) and np.all(got[total:] == 0x5A5A5A5A)
            else:
                assert np.all(got == 0x5A5A5A5A)
        off = torch.empty(nq + 1, dtype=torch.int64, device=dev)
        t = C.c_uint64(0)
        assert lib.lapper64_find_batch_dev(h, vp(tq), vp(tqe), nq, vp(off), C.c_void_p(0), 0, C.byref(t), sp) == _ffi.OK
        assert int(t.value) == total
    finally:
        lib.lapper64_free(h)


# ---------------------------------------------------------------------------------------------- narrowing
# A u64 set whose coordinates fit 32 bits is answered by the u32 engine (csrc/engine64.cu, "narrowing"): same oracle
# (I = u64), same bits, whichever engine runs.
@pytest.fixture
def narrow_mode():
    from rust_lapper_b200 import _ffi

    lib = _ffi.lib()
    old = []

    def set_mode(m):
        old.append(lib.lapper64_debug_set_narrow_mode(m))

    yield set_mode
    if old:
        lib.lapper64_debug_set_narrow_mode(old[0])


@pytest.mark.parametrize("mode", [0, 2], ids=["wide", "narrow"])
@pytest.mark.parametrize("check", rs.CORE_CHECKS + [rs.check_depth], ids=lambda f: f.__name__)
def test_reference_known_answers_u64_either_engine(make, narrow_mode, mode, check):
    narrow_mode(mode)
    check(make)


def _special_queries(rng, universe, nq, qlen):
    qs = rng.integers(0, universe + universe // 8, nq, dtype=np.uint64)
    qe = qs + rng.integers(0, max(qlen, 2), nq, dtype=np.uint64)
    k = rng.choice(nq, nq // 10, replace=False)
    qe[k] = qs[k] - rng.integers(0, 3, k.size, dtype=np.uint64).clip(max=qs[k])  # empty and inverted
    top32 = (1 << 32) - 1
    pts = np.array([0, 1, top32 - 3, top32 - 2, top32 - 1, top32, top32 + 1, 1 << 40, U64_MAX - 1, U64_MAX], dtype=np.uint64)
    a, b = np.meshgrid(pts, pts)
    m = a.size
    qs[:m], qe[:m] = a.reshape(-1), b.reshape(-1)
    return qs, qe


@pytest.mark.parametrize("n,universe,max_len,top", [
    (1, 1000, 50, None),
    (3000, 200_000, 300, None),
    (200_000, 3_000_000_000, 5000, None),
    (200_000, 3_000_000_000, 5000, (1 << 32) - 3),   # an interval that ends at the largest coordinate that still narrows
    (50_000, 1 << 20, 1 << 8, None),                  # duplicates and ties
])
def test_narrow_engine_matches_the_u64_oracle(oracle64, narrow_mode, n, universe, max_len, top):
    from rust_lapper_b200 import Lapper64

    narrow_mode(2)
    rng = np.random.default_rng(n ^ 0x5EED)
    s, e = _random_set(rng, n, universe, max_len, 0.0)
    if top is not None:
        s[0], e[0] = top - 7, top
    gpu, cpu = Lapper64(starts=s, stops=e), oracle64(starts=s, stops=e)
    qs, qe = _special_queries(rng, universe, 30_000, max_len)
    _same(gpu, cpu, qs, qe)
    assert gpu.is_narrow
    o = np.argsort(qs, kind="stable")
    _same(gpu, cpu, qs[o], np.maximum.accumulate(qe[o]))
    # the set operations and the structure still come from the u64 arrays
    gs, ge, gp = gpu.intervals_soa()
    cs, ce, cv = cpu.intervals_soa()
    assert np.array_equal(gs, cs) and np.array_equal(ge, ce) and np.array_equal(gp, cv)
    assert gpu.cov() == cpu.cov()
    # merge_overlaps drops the u32 index; the next batch rebuilds it over the merged set
    gpu.merge_overlaps(), cpu.merge_overlaps()
    assert not gpu.is_narrow
    _same(gpu, cpu, qs, qe)
    assert gpu.is_narrow
    # insert beyond 32 bits: the handle goes back to the 64-bit kernels for good
    gpu.insert(1 << 33, (1 << 33) + 10), cpu.insert(1 << 33, (1 << 33) + 10)
    _same(gpu, cpu, qs, qe)
    assert not gpu.is_narrow


def test_narrow_boundary_and_default_policy(oracle64, narrow_mode):
    from rust_lapper_b200 import Lapper64

    rng = np.random.default_rng(99)
    s, e = _random_set(rng, 5000, 1_000_000, 400, 0.0)
    qs, qe = _special_queries(rng, 1_000_000, 30_000, 400)
    # default policy: small batches stay on the 64-bit kernels, the first batch of >= 4096 queries builds the u32 index
    gpu, cpu = Lapper64(starts=s, stops=e), oracle64(starts=s, stops=e)
    _same(gpu, cpu, qs[:1000], qe[:1000])
    assert not gpu.is_narrow
    _same(gpu, cpu, qs, qe)
    assert gpu.is_narrow
    _same(gpu, cpu, qs[:1000], qe[:1000])  # small batches use it once it exists
    # a set that starts at 0 with one coordinate above 2^32 - 3: never narrowed (b' = 2^32 - 2 could no longer stand in
    # for every larger stop, and the span is too wide for the rebased form)
    narrow_mode(2)
    for top, narrows in (((1 << 32) - 3, True), ((1 << 32) - 2, False), ((1 << 32) - 1, False), (1 << 32, False)):
        s2, e2 = s.copy(), e.copy()
        s2[0], e2[0] = top - 3, top
        s2[1], e2[1] = 0, 5
        gpu, cpu = Lapper64(starts=s2, stops=e2), oracle64(starts=s2, stops=e2)
        _same(gpu, cpu, qs, qe)
        assert gpu.is_narrow == narrows, top
    # inverted intervals narrow too (count / find have no precondition)
    s3, e3 = s.copy(), e.copy()
    s3[:50], e3[:50] = e[:50] + np.uint64(5), s[:50]
    gpu, cpu = Lapper64(starts=s3, stops=e3), oracle64(starts=s3, stops=e3)
    assert gpu.has_inverted
    _same(gpu, cpu, qs, qe)
    assert gpu.is_narrow


@pytest.mark.parametrize("place", ["2^40", "2^63", "top"])
@pytest.mark.parametrize("span,narrows", [(5_000_000, True), ((1 << 32) - 4, True), ((1 << 32) - 3, False)])
def test_rebased_narrowing_matches_the_u64_oracle(oracle64, narrow_mode, place, span, narrows):
    """Coordinates beyond 32 bits whose span fits: the u32 engine answers on coordinates relative to the smallest one
    (csrc/engine64.cu, "rebased form"); one coordinate more of span and the 64-bit kernels answer.  Same bits either way."""
    from rust_lapper_b200 import Lapper64

    narrow_mode(2)
    rng = np.random.default_rng(span % 1000 + len(place))
    base = {"2^40": 1 << 40, "2^63": (1 << 63) - 12345, "top": U64_MAX - span}[place]
    n = 60_000
    rel_s = rng.integers(0, span - 3010, n, dtype=np.uint64)
    rel_e = rel_s + rng.integers(0, 3000, n, dtype=np.uint64)
    rel_s[0], rel_e[0] = 0, 7                 # the smallest coordinate: the base of the map
    rel_s[1], rel_e[1] = span - 9, span       # the largest one
    rel_s[2:40], rel_e[2:40] = rel_e[2:40] + np.uint64(3), rel_s[2:40].copy()   # a few inverted intervals
    s, e = rel_s + np.uint64(base), rel_e + np.uint64(base)
    gpu, cpu = Lapper64(starts=s, stops=e), oracle64(starts=s, stops=e)
    nq = 40_000
    qs = np.uint64(base) + rng.integers(0, span, nq, dtype=np.uint64)
    qe = np.minimum(qs.astype(object) + rng.integers(0, 5000, nq).astype(object), U64_MAX).astype(np.uint64)
    k = rng.choice(nq, nq // 10, replace=False)
    qe[k] = qs[k] - rng.integers(0, 3, k.size, dtype=np.uint64)   # empty and inverted queries
    hi = base + span
    pts = np.array([0, 1, (1 << 32) - 1, base - 1, base, base + 1, base + 7, hi - 9, hi - 1, hi, min(hi + 1, U64_MAX),
                    min(hi + (1 << 33), U64_MAX), U64_MAX - 1, U64_MAX], dtype=np.uint64)
    a, b = np.meshgrid(pts, pts)
    qs[:a.size], qe[:a.size] = a.reshape(-1), b.reshape(-1)
    _same(gpu, cpu, qs, qe)
    assert gpu.is_narrow == narrows
    o = np.argsort(qs, kind="stable")
    _same(gpu, cpu, qs[o], np.maximum.accumulate(qe[o]))
    # structure and set operations come from the u64 arrays
    assert gpu.max_len == cpu.max_len and np.array_equal(gpu.stops, cpu.stops)
    # an insert below the base moves it: the u32 index is dropped and rebuilt on the new base (or not at all)
    lo = base - 1000
    gpu.insert(lo, lo + 10), cpu.insert(lo, lo + 10)
    assert not gpu.is_narrow
    _same(gpu, cpu, qs, qe)
    assert gpu.is_narrow == (span + 1000 <= (1 << 32) - 4)


def test_narrow_one_call_device_api(oracle64, narrow_mode):
    """lapper64_find_batch_dev / lapper64_count_batch_dev on a narrowed handle: same capacity protocol, same bits"""
    import ctypes as C

    import torch

    from rust_lapper_b200 import Lapper64, _ffi

    narrow_mode(2)
    lib = _ffi.lib()
    rng = np.random.default_rng(5)
    s, e = _random_set(rng, 100_000, 50_000_000, 2000, 0.0)
    gpu, cpu = Lapper64(starts=s, stops=e), oracle64(starts=s, stops=e)
    qs, qe = _special_queries(rng, 50_000_000, 100_000, 500)
    wo, wi = cpu.find_batch(qs, qe, 4)
    total = int(wo[-1])
    dev = torch.device("cuda", 0)
    dq_s = torch.from_numpy(qs.view(np.int64)).to(dev)
    dq_e = torch.from_numpy(qe.view(np.int64)).to(dev)
    d_off = torch.zeros(qs.size + 1, dtype=torch.int64, device=dev)
    d_cnt = torch.zeros(qs.size, dtype=torch.int64, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _ffi.check(lib.lapper64_count_batch_dev(gpu._h, vp(dq_s), vp(dq_e), qs.size, vp(d_cnt), sp))
    torch.cuda.synchronize()
    assert np.array_equal(d_cnt.cpu().numpy().view(np.uint64), cpu.count_batch(qs, qe, 4))
    tot = C.c_uint64(0)
    # too small a buffer: LAPPER_ERR_CAPACITY with complete offsets and total
    d_idx = torch.zeros(max(total // 2, 1), dtype=torch.int32, device=dev)
    rc = lib.lapper64_find_batch_dev(gpu._h, vp(dq_s), vp(dq_e), qs.size, vp(d_off), vp(d_idx), d_idx.numel(), C.byref(tot), sp)
    assert rc == _ffi.ERR_CAPACITY and int(tot.value) == total
    assert np.array_equal(d_off.cpu().numpy().view(np.uint64), wo)
    d_idx = torch.zeros(total + 5, dtype=torch.int32, device=dev)
    _ffi.check(lib.lapper64_find_batch_dev(gpu._h, vp(dq_s), vp(dq_e), qs.size, vp(d_off), vp(d_idx), d_idx.numel(), C.byref(tot), sp))
    assert int(tot.value) == total
    assert np.array_equal(d_off.cpu().numpy().view(np.uint64), wo)
    assert np.array_equal(d_idx[:total].cpu().numpy().view(np.uint32), wi)
    assert gpu.is_narrow


@pytest.mark.parametrize("mode", [0, 1], ids=["wide", "default"])
def test_host_count_in_several_chunks_u64(oracle64, narrow_mode, mode):
    """lapper64_count_batch pipelines batches beyond 4 M queries in chunks (H2D | count | D2H on three streams)"""
    from rust_lapper_b200 import Lapper64

    narrow_mode(mode)
    rng = np.random.default_rng(31)
    sh = np.uint64(0 if mode else 1 << 36)
    s, e = _random_set(rng, 100_000, 50_000_000, 3000, 0.0)
    s, e = s + sh, e + sh
    gpu, cpu = Lapper64(starts=s, stops=e), oracle64(starts=s, stops=e)
    nq = 9_000_001
    qs = rng.integers(0, 55_000_000, nq, dtype=np.uint64) + sh
    qe = qs + rng.integers(0, 600, nq, dtype=np.uint64)
    qs[-3:] = [0, U64_MAX, 1 << 32]
    qe[-3:] = [U64_MAX, U64_MAX, 1 << 33]
    assert np.array_equal(gpu.count_batch(qs, qe), cpu.count_batch(qs, qe, 8))
    assert gpu.is_narrow == bool(mode)
