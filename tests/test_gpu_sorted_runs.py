"""Parity of the sorted-run kernel (csrc/count_sorted.cu) with the oracle.  The kernel is offered to large batches
and selected on the device by the batch probe; the test hook lowers "large" so that oracle-sized batches reach it.
Covered: clean sorted dense batches (count u64 / u32, find CSR), runs that do not qualify inside a batch that does
(shuffled runs, a == b and inverted queries, queries ending at u32::MAX, clusters of more than 128 intervals under one
run), duplicates, the ragged tail, interval sets with inverted intervals (count only) -- always bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RUN = 256


@pytest.fixture(autouse=True)
def _offer_sorted_kernel_to_small_batches():
    from rust_lapper_b200 import _ffi

    lib = _ffi.lib()
    prev = lib.lapper_debug_set_sorted_min_queries(32 * 1024)
    yield
    lib.lapper_debug_set_sorted_min_queries(prev)


def _same(gpu, cpu, qs, qe, find=True):
    want = cpu.count_batch(qs, qe)
    assert np.array_equal(gpu.count_batch(qs, qe), want)
    assert np.array_equal(gpu.count_batch_c32(qs, qe), want.astype(np.uint32))
    if find:
        go, gi = gpu.find_batch(qs, qe)
        co, ci = cpu.find_batch(qs, qe)
        assert np.array_equal(go, co)
        assert np.array_equal(gi, ci)


def _launches(fn):
    from rust_lapper_b200 import _ffi

    lib = _ffi.lib()
    n0 = int(lib.lapper_query_kernel_launches())
    fn()
    return int(lib.lapper_query_kernel_launches()) - n0


def _run_stats(reset=True):
    """(runs answered by ranking, runs answered per query) since the last reset"""
    import ctypes as C

    from rust_lapper_b200 import _ffi

    out = (C.c_uint64 * 2)()
    _ffi.check(_ffi.lib().lapper_debug_sorted_run_stats(out, 1 if reset else 0))
    return int(out[0]), int(out[1])


def _sets(oracle, s, e, flags=0):
    from rust_lapper_b200 import Lapper

    return Lapper(starts=s, stops=e, flags=flags), oracle.OracleLapper(starts=s, stops=e)


@pytest.mark.parametrize("shape", ["uniform", "gene_exon"])
@pytest.mark.parametrize("nq", [64 * 1024, 200_000 + 77])
def test_clean_sorted_batch(oracle, shape, nq):
    from rust_lapper_b200 import synth

    U = 40_000_000
    n = 8_000  # the probe wants <= 40 expected intervals per run of 256 queries: >= ~7 queries per interval
    if shape == "uniform":
        s, e = synth.intervals_uniform(42, n, U, 50, 5000)
    else:
        s, e = synth.intervals_gene_exon(44, n, U)
    gpu, cpu = _sets(oracle, s, e)
    assert gpu.bucket_shift is not None
    qs, qe = synth.queries_sorted_fixed_len(45, nq, U, 150)
    _run_stats()
    _same(gpu, cpu, qs, qe)
    fast, slow = _run_stats()
    # count (u64), count (u32) and the offsets pass of find each went through the kernel: nearly every run by ranking
    assert fast + slow == 3 * (nq // 1024) * 4
    assert fast >= 0.9 * (fast + slow)


def test_sorted_kernel_is_launched_for_sorted_batches_only(oracle):
    from rust_lapper_b200 import _ffi, synth

    U = 40_000_000
    s, e = synth.intervals_uniform(42, 8_000, U, 50, 5000)
    gpu, cpu = _sets(oracle, s, e)
    qs, qe = synth.queries_sorted_fixed_len(45, 64 * 1024, U, 150)
    lib = _ffi.lib()
    with_hook = _launches(lambda: gpu.count_batch_c32(qs, qe))
    prev = lib.lapper_debug_set_sorted_min_queries(1 << 40)
    without = _launches(lambda: gpu.count_batch_c32(qs, qe))
    lib.lapper_debug_set_sorted_min_queries(prev)
    assert with_hook == without + 1
    # ... and it exits at once for a batch that is not sorted: same answers either way
    rq = synth.queries_fixed_len(43, 64 * 1024, U, 150)
    assert np.array_equal(gpu.count_batch(*rq), cpu.count_batch(*rq))


def test_variable_lengths_duplicates_and_ties(oracle):
    """starts sorted, stops sorted too (lengths grow slowly), many equal queries and equal interval coordinates"""
    rng = np.random.default_rng(5)
    U = 3_000_000
    n = 9_000
    s = (rng.integers(0, U // 64, n) * 64).astype(np.uint32)  # many ties
    e = (s + (rng.integers(0, 40, n) * 16)).astype(np.uint32)  # incl. empty intervals
    gpu, cpu = _sets(oracle, s, e)
    nq = 96 * 1024 + 5
    qs = np.sort(rng.integers(0, U, nq)).astype(np.int64)
    qs[1000:1600] = qs[1000]                                   # a long stretch of identical queries
    qs = np.sort(qs)
    ln = 100 + (np.arange(nq) // 4096)  # non-decreasing lengths
    qe = qs + ln
    qe = np.maximum.accumulate(qe)
    _same(gpu, cpu, qs.astype(np.uint32), qe.astype(np.uint32))


def test_runs_that_do_not_qualify(oracle):
    """a batch the probe accepts (its sampled runs are clean) with dirty runs in between"""
    from rust_lapper_b200 import synth

    rng = np.random.default_rng(9)
    U = 30_000_000
    s, e = synth.intervals_uniform(42, 12_000, U, 50, 5000)
    # a cluster of 3000 intervals inside a few hundred coordinates: more than 128 starts and stops under one run
    cs = (7_000_000 + rng.integers(0, 300, 3000)).astype(np.uint32)
    ce = (cs + rng.integers(1, 200, 3000)).astype(np.uint32)
    s, e = np.concatenate([s, cs]), np.concatenate([e, ce])
    gpu, cpu = _sets(oracle, s, e)
    nq = 128 * 1024 + 300
    qs, qe = synth.queries_sorted_fixed_len(45, nq, U, 150)
    qs, qe = qs.astype(np.int64), qe.astype(np.int64)
    tiles = nq // 1024
    sampled = {(tiles * lane // 32) * 4 for lane in range(32)}  # runs the probe looks at (first run of 32 tiles)
    runs = [r for r in range(nq // RUN) if r not in sampled]
    pick = rng.choice(runs, 40, replace=False)
    for i, r in enumerate(pick):
        lo = r * RUN
        kind = i % 5
        if kind == 0:    # shuffled run
            p = rng.permutation(RUN)
            qs[lo:lo + RUN], qe[lo:lo + RUN] = qs[lo:lo + RUN][p], qe[lo:lo + RUN][p]
        elif kind == 1:  # starts sorted, stops not
            qe[lo + 17] += 5000
        elif kind == 2:  # empty and inverted queries
            qe[lo + 3] = qs[lo + 3]
            qe[lo + 200] = max(qs[lo + 200] - 40, 0)
        elif kind == 3:  # the run's last queries at u32::MAX
            qs[lo + RUN - 1] = 0xFFFFFFFF
            qe[lo + RUN - 1] = 0xFFFFFFFF
            qs[lo + RUN - 2] = 0xFFFFFFFE
            qe[lo + RUN - 2] = 0xFFFFFFFF
        else:            # a long query in the middle (stops out of order)
            qe[lo + 100] = min(qs[lo + 100] + 2_000_000, 0xFFFFFFFF)
    _run_stats()
    _same(gpu, cpu, qs.astype(np.uint32), qe.astype(np.uint32))
    fast, slow = _run_stats()
    assert fast > 0 and slow >= 3 * 30  # the dirty runs (and the runs over the cluster) were answered per query


def test_sorted_batch_over_inverted_intervals(oracle):
    """count has no precondition on the intervals (src/lib.rs:610-618); find takes the general path there"""
    from rust_lapper_b200 import synth

    rng = np.random.default_rng(13)
    U = 20_000_000
    s, e = synth.intervals_uniform(42, 8_000, U, 50, 5000)
    s, e = s.copy(), e.copy()
    k = rng.choice(s.size, 300, replace=False)
    s[k], e[k] = e[k], s[k]
    gpu, cpu = _sets(oracle, s, e)
    assert gpu.has_inverted
    qs, qe = synth.queries_sorted_fixed_len(45, 70_000, U, 150)
    _same(gpu, cpu, qs, qe)


def test_sorted_batch_at_the_ends_of_the_coordinate_range(oracle):
    """runs below the first interval, beyond the last one, and up to u32::MAX (the sentinel sector; ranks 0 and n)"""
    rng = np.random.default_rng(21)
    n = 6_000
    s = (1_000_000 + np.sort(rng.integers(0, 10_000_000, n))).astype(np.uint32)
    e = (s + rng.integers(1, 3000, n)).astype(np.uint32)
    gpu, cpu = _sets(oracle, s, e)
    nq = 64 * 1024
    a = np.sort(rng.integers(0, 14_000_000, nq - 4096)).astype(np.int64)
    hi = np.sort(rng.integers(0xFFFFFFFF - 2_000_000, 0xFFFFFFFF - 200, 4096)).astype(np.int64)
    qs = np.concatenate([a, hi])
    qe = qs + 150
    _same(gpu, cpu, qs.astype(np.uint32), qe.astype(np.uint32))


@pytest.mark.parametrize("flags", [0, 8 | 2])
def test_dense_set_sorted_queries(oracle, flags):
    """more intervals than the run kernel takes per run in places: mixed qualifying / per-query runs"""
    from rust_lapper_b200 import synth

    U = 6_000_000
    s, e = synth.intervals_gene_exon(44, 40_000, U)  # ~ 40 intervals per run on average, clustered
    gpu, cpu = _sets(oracle, s, e, flags)
    qs, qe = synth.queries_sorted_fixed_len(45, 250_000, U, 100)
    _same(gpu, cpu, qs, qe)
