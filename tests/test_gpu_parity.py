"""Parity of the CUDA path (through the C ABI) with the CPU oracle on identical seeded inputs:
bit-exact counts, CSR offsets and positions; merge / cov / union_and_intersect; seek cursors."""
import numpy as np
import pytest

import reference_suite as rs

pytestmark = pytest.mark.gpu

FLAGS = {"auto": 0, "plain": 1, "shift4": (4 + 1) << 8, "shift13": (13 + 1) << 8}


def _both(oracle, s, e, flags=0):
    from rust_lapper_b200 import Lapper

    return Lapper(starts=s, stops=e, flags=flags), oracle.OracleLapper(starts=s, stops=e)


def _assert_same_queries(gpu, cpu, qs, qe):
    assert np.array_equal(gpu.count_batch(qs, qe), cpu.count_batch(qs, qe))
    assert np.array_equal(gpu.count_batch_c32(qs, qe), cpu.count_batch(qs, qe).astype(np.uint32))
    go, gi = gpu.find_batch(qs, qe)
    co, ci = cpu.find_batch(qs, qe)
    assert np.array_equal(go, co)
    assert np.array_equal(gi, ci)


@pytest.mark.parametrize("flags", list(FLAGS))
@pytest.mark.parametrize("seed", range(3))
def test_random_small(oracle, seed, flags):
    rng = np.random.default_rng(seed)
    for _ in range(12):
        n = int(rng.integers(1, 400))
        universe = int(rng.choice([50, 3000, 10**6, 4 * 10**9]))
        max_len = int(rng.choice([0, 5, 200, universe // 3]))
        nq = int(rng.integers(1, 700))
        s, e, qs, qe = rs.random_case(rng, n, min(universe, 0xFFFFFFFF - 1), max_len, nq,
                                      int(rng.choice([0, 1, 100, universe // 2])))
        gpu, cpu = _both(oracle, s, e, FLAGS[flags])
        assert len(gpu) == len(cpu) and gpu.max_len == cpu.max_len
        gs, ge, gp = gpu.intervals_soa()
        cs, ce, cv = cpu.intervals_soa()
        assert np.array_equal(gs, cs) and np.array_equal(ge, ce) and np.array_equal(gp, cv)
        assert np.array_equal(gpu.starts, cpu.starts) and np.array_equal(gpu.stops, cpu.stops)
        _assert_same_queries(gpu, cpu, qs, qe)


@pytest.mark.parametrize("n", [2, 127, 128, 129, 4095, 4096, 4097, 9000])
def test_sorted_structure_around_the_small_sort_limit(oracle, n):
    """csrc/build.cu sorts sets of <= 4096 intervals by rank counting in shared memory (128 keys per CTA) and larger ones
    by radix passes: both stable (src/lib.rs:201), so the sorted vector, `perm` and the private `stops` are the oracle's on
    either side of the limit -- with many duplicates of (start, stop) and of start alone, and inverted intervals."""
    rng = np.random.default_rng(n)
    s = rng.integers(0, max(n // 6, 2), n).astype(np.uint32) * np.uint32(1000)
    e = s + rng.integers(0, 3, n).astype(np.uint32) * np.uint32(700)
    s[::5], e[::5] = 0xFFFFFF00, 0xFFFFFFFF                    # one large key many times over
    k = rng.random(n) < 0.05
    s[k], e[k] = e[k].copy(), s[k].copy()                      # inverted where e != s
    gpu, cpu = _both(oracle, s, e, FLAGS["auto"])
    gs, ge, gp = gpu.intervals_soa()
    cs, ce, cv = cpu.intervals_soa()
    assert np.array_equal(gs, cs) and np.array_equal(ge, ce) and np.array_equal(gp, cv)
    assert np.array_equal(gpu.starts, cpu.starts) and np.array_equal(gpu.stops, cpu.stops) and gpu.max_len == cpu.max_len
    near = s[rng.integers(0, n, 2000)].astype(np.int64) + rng.integers(-1500, 1500, 2000)
    qs = np.concatenate([np.clip(near, 0, 0xFFFFFFFF), rng.integers(0, 0xFFFFFFFF, 1000)]).astype(np.uint32)
    _assert_same_queries(gpu, cpu, qs, qs + np.uint32(900))


@pytest.mark.parametrize("flags", ["auto", "plain"])
def test_inverted_intervals_and_queries(oracle, flags):
    rng = np.random.default_rng(11)
    n, nq = 300, 2000
    s = rng.integers(0, 5000, n).astype(np.uint32)
    e = rng.integers(0, 5000, n).astype(np.uint32)  # about half inverted
    qs = rng.integers(0, 5200, nq).astype(np.uint32)
    qe = rng.integers(0, 5200, nq).astype(np.uint32)
    gpu, cpu = _both(oracle, s, e, FLAGS[flags])
    assert gpu.max_len == cpu.max_len
    _assert_same_queries(gpu, cpu, qs, qe)


# ---- the packed table in front of the sector table (csrc/packed_sector.cuh).  It is only consulted by full,
# aligned tiles of 1024 queries whose warps see spread-out queries, so these cases use >= 1024 random queries.
PACKED_FLAGS = {"auto_width": 2 | 8, "w2": 2 | (2 << 16), "w7": 2 | (7 << 16), "w100": 2 | (100 << 16),
                "w257": 2 | (257 << 16), "w1550": 2 | (1550 << 16), "w1820": 2 | (1820 << 16),
                "w1000_shift3": (1000 << 16) | ((3 + 1) << 8)}


@pytest.mark.parametrize("flags", list(PACKED_FLAGS))
def test_packed_table_random(oracle, flags):
    rng = np.random.default_rng(100 + list(PACKED_FLAGS).index(flags))
    for universe, n, max_len, qlen in [(3_000_000, 20_000, 700, 150), (60_000, 20_000, 10, 3), (4 * 10**9, 50_000, 5000, 150),
                                       (3_000_000, 3_000, 400_000, 30), (5_000_000, 30_000, 0, 2500),
                                       (1_200_000, 20_000, 3000, 150)]:  # last: buckets narrower than the queries (minimum margin)
        s, e, qs, qe = rs.random_case(rng, n, min(universe, 0xFFFFFFFF - 1), max_len, 6 * 1024 + 77, qlen)
        gpu, cpu = _both(oracle, s, e, PACKED_FLAGS[flags])
        assert gpu.packed_table is not None
        _assert_same_queries(gpu, cpu, qs, qe)
        # mixed batch: a quarter of the queries long / inverted / at the extremes -> per-query fallback inside packed tiles
        k = qs.size // 4
        qs2, qe2 = qs.copy(), qe.copy()
        idx = rng.permutation(qs.size)[:k]
        qe2[idx[: k // 3]] = np.minimum(qs2[idx[: k // 3]].astype(np.uint64) + universe // 3, 0xFFFFFFFF).astype(np.uint32)
        qs2[idx[k // 3: 2 * k // 3]], qe2[idx[k // 3: 2 * k // 3]] = qe2[idx[k // 3: 2 * k // 3]], qs2[idx[k // 3: 2 * k // 3]].copy()
        qs2[idx[2 * k // 3:]] = 0xFFFFFFFF
        _assert_same_queries(gpu, cpu, qs2, qe2)


def test_packed_table_is_built_and_used_where_it_pays(oracle):
    """Auto mode: the packed table exists only next to a sector table too large for L2 (none at this size),
    and forcing it changes nothing about the answers, on sorted (sector-table path) or random queries."""
    from rust_lapper_b200 import Lapper, synth

    U = synth.GRCH38 // 20
    s, e = synth.intervals_uniform(42, 500_000, U, 50, 5000)
    qs, qe = synth.queries_fixed_len(43, 1_000_000, U, 150)
    auto = Lapper(starts=s, stops=e)
    assert auto.packed_table is None
    forced = Lapper(starts=s, stops=e, flags=8)
    w, nsect, nover = forced.packed_table
    assert 16 <= w <= 1820 and nsect == (int(max(s.max(), e.max())) // w) + 2 and nover < nsect // 8
    want = auto.count_batch(qs, qe)
    assert np.array_equal(forced.count_batch(qs, qe), want)
    order = np.argsort(qs, kind="stable")
    assert np.array_equal(forced.count_batch(qs[order], qe[order]), want[order])
    cpu = oracle.OracleLapper(starts=s, stops=e)
    assert np.array_equal(want[:200_000], cpu.count_batch(qs[:200_000], qe[:200_000]))
    go, gi = forced.find_batch(qs[:300_000], qe[:300_000])
    co, ci = cpu.find_batch(qs[:300_000], qe[:300_000])
    assert np.array_equal(go, co) and np.array_equal(gi, ci)


@pytest.mark.parametrize("depth", [12, 40, 150, 260, 600])
@pytest.mark.parametrize("order", ["sorted", "random"])
def test_find_tile_emit_dense_groups(oracle, depth, order):
    """The tile-form emit kernel (find_fill_tile_kernel) away from its comfortable case: groups of 64 queries whose
    hits exceed the warp's stage (sub-range passes), more than 64 candidates per group (per-lane merge into the stage),
    queries around the heavy threshold of 256 hits (handed to find_heavy_kernel), several length classes, ragged last
    tile; and the same queries in random order (tiles without locality fall back to the per-warp kernel)."""
    rng = np.random.default_rng(depth)
    U = 400_000
    n = 12_000
    mean_len = depth * U // n
    ln = np.concatenate([rng.integers(1, 2 * mean_len, n - n // 8), rng.integers(20 * mean_len, 40 * mean_len, n // 8) // 16])
    s = rng.integers(0, U, n)
    e = np.minimum(s + ln, 0xFFFFFFFF)
    s, e = s.astype(np.uint32), e.astype(np.uint32)
    nq = 7 * 1024 + 333
    qs = rng.integers(0, U + 1000, nq)
    if order == "sorted":
        qs = np.sort(qs)
    qe = qs + rng.integers(0, 300, nq)
    qs, qe = qs.astype(np.uint32), qe.astype(np.uint32)
    gpu, cpu = _both(oracle, s, e)
    go, gi = gpu.find_batch(qs, qe)
    co, ci = cpu.find_batch(qs, qe)
    assert np.array_equal(go, co)
    assert np.array_equal(gi, ci)
    hits = np.diff(co)
    assert hits.max() > 2 * depth // 3  # the shape is what the name says


@pytest.mark.parametrize("shape", ["one_long_class", "three_long_classes", "long_and_dense_short"])
def test_heavy_queries_sparse_windows(oracle, shape):
    """Heavy queries (>= 256 hits) whose hits are a small fraction of the sorted vector's window: find_heavy_kernel
    scans the length-class windows and merges their hit streams by position (heavy_merge_classes) -- one, several
    and unevenly sized classes, duplicates and ties included."""
    rng = np.random.default_rng({"one_long_class": 1, "three_long_classes": 2, "long_and_dense_short": 3}[shape])
    U = 50_000_000
    n = 300_000
    s = rng.integers(0, U, n)
    ln = rng.integers(10, 200, n)
    if shape == "one_long_class":
        k = n // 100
        ln[:k] = rng.integers(U // 2, 9 * U // 10, k)
    elif shape == "three_long_classes":
        k = n // 150
        ln[:k] = rng.integers(U // 2, 9 * U // 10, k)
        ln[k:2 * k] = rng.integers(U // 40, U // 20, k)
        ln[2 * k:3 * k] = rng.integers(U // 400, U // 200, k)
        s[3 * k:3 * k + 500] = s[0]  # ties on start across classes
    else:
        k = n // 200
        ln[:k] = rng.integers(U // 3, U // 2, k)
        s[k:k + 40_000] = rng.integers(U // 2, U // 2 + 3000, 40_000)  # a dense cluster of short intervals
    e = np.minimum(s + ln, 0xFFFFFFFF)
    s, e = s.astype(np.uint32), e.astype(np.uint32)
    nq = 1024 + 40
    qs = rng.integers(U // 4, 3 * U // 4, nq)
    qs[:16] = U // 2 + rng.integers(0, 3000, 16)  # into the dense cluster
    qe = qs + rng.integers(1, 400, nq)
    qs, qe = qs.astype(np.uint32), qe.astype(np.uint32)
    gpu, cpu = _both(oracle, s, e)
    go, gi = gpu.find_batch(qs, qe)
    co, ci = cpu.find_batch(qs, qe)
    assert np.array_equal(go, co)
    assert np.array_equal(gi, ci)
    assert np.diff(co).max() >= 256


@pytest.mark.parametrize("length", [0, 1, 37, 1000])
def test_find_tile_emit_single_length_class(oracle, length):
    """All intervals of one length (one length class; length 0 and 1 are never / barely found): start-sorted queries
    through the tile emit kernel, ragged last tile, unaligned `indices` slices."""
    rng = np.random.default_rng(1000 + length)
    U = 300_000
    s = rng.integers(0, U, 40_000)
    e = s + length
    nq = 5 * 1024 + 17
    qs = np.sort(rng.integers(0, U, nq))
    qe = qs + rng.integers(0, 120, nq)
    gpu, cpu = _both(oracle, s.astype(np.uint32), e.astype(np.uint32))
    go, gi = gpu.find_batch(qs.astype(np.uint32), qe.astype(np.uint32))
    co, ci = cpu.find_batch(qs.astype(np.uint32), qe.astype(np.uint32))
    assert np.array_equal(go, co)
    assert np.array_equal(gi, ci)


def test_extreme_coordinates(oracle):
    X = 0xFFFFFFFF
    s = np.array([0, 0, X - 5, X, X - 1, 7, X], dtype=np.uint32)
    e = np.array([X, 0, X, X, X, 7, X - 1], dtype=np.uint32)
    vals = [0, 1, 5, 7, X - 6, X - 5, X - 1, X]
    qs, qe = np.meshgrid(vals, vals)
    qs, qe = qs.ravel().astype(np.uint32), qe.ravel().astype(np.uint32)
    for flags in (0, 1, 2, 2 | ((13 + 1) << 8), 2 | 8, 2 | (1820 << 16), 2 | (3 << 16)):
        gpu, cpu = _both(oracle, s, e, flags)
        _assert_same_queries(gpu, cpu, qs, qe)
        # the same 64 queries tiled to full 1024-query tiles (the packed path only sees full tiles)
        _assert_same_queries(gpu, cpu, np.tile(qs, 40), np.tile(qe, 40))


def test_empty_and_ragged(oracle):
    from rust_lapper_b200 import Lapper

    gpu = Lapper([])
    assert len(gpu) == 0 and gpu.max_len == 0
    assert gpu.count_batch([1, 2, 3], [4, 5, 6]).tolist() == [0, 0, 0]
    off, idx = gpu.find_batch([1, 2, 3], [4, 5, 6])
    assert off.tolist() == [0, 0, 0, 0] and idx.size == 0
    one = Lapper([(10, 20)])
    assert one.count_batch([], []).size == 0
    off, idx = one.find_batch([], [])
    assert off.tolist() == [0] and idx.size == 0
    # ragged batch sizes around the 4-query / 1024-query tiles
    rng = np.random.default_rng(5)
    s, e, _, _ = rs.random_case(rng, 5000, 10**6, 3000, 1, 1)
    gpu, cpu = _both(oracle, s, e)
    for nq in (1, 2, 3, 4, 5, 255, 1023, 1024, 1025, 4099):
        _, _, qs, qe = rs.random_case(rng, 1, 10**6, 3000, nq, 500)
        _assert_same_queries(gpu, cpu, qs, qe)
        # unaligned device pointers take the scalar path: offset the host view by one element
        _assert_same_queries(gpu, cpu, qs[1:], qe[1:])


@pytest.mark.parametrize("name", ["cfg2_uniform", "cfg3_gene_exon", "cfg4_pathological", "cfg1b_criterion", "dense"])
def test_seeded_configs_medium(oracle, name):
    """The BASELINE.json shapes at sizes the oracle finishes in seconds (same generators as bench.py)."""
    from rust_lapper_b200 import synth

    U = synth.GRCH38
    if name == "cfg2_uniform":
        s, e = synth.intervals_uniform(42, 200_000, U // 50, 50, 5000)
        qs, qe = synth.queries_fixed_len(43, 300_000, U // 50, 150)
    elif name == "cfg3_gene_exon":
        s, e = synth.intervals_gene_exon(44, 200_000, U // 50)
        qs, qe = synth.queries_sorted_fixed_len(45, 300_000, U // 50, 150)
    elif name == "cfg4_pathological":
        s, e = synth.intervals_pathological(46, 20_000, U)
        qs, qe = synth.queries_fixed_len(47, 2_000, U, 150)
    elif name == "cfg1b_criterion":  # benches/lapper_benchmark.rs:34-43: n = 1000, U = 1e8, len 500..80000
        s, e = synth.intervals_uniform(7, 1000, 100_000_000, 500, 79_999)
        qs, qe = synth.queries_fixed_len(8, 1000, 1_000_000_000, 1)
        qs, qe = np.concatenate([qs, s]), np.concatenate([qe, e])
    else:  # far more keys than coordinates: every bucket overflows
        s, e = synth.intervals_uniform(9, 100_000, 3000, 0, 40)
        qs, qe = synth.queries_fixed_len(10, 50_000, 3100, 7)
    gpu, cpu = _both(oracle, s, e)
    _assert_same_queries(gpu, cpu, qs, qe)
    assert gpu.cov() == cpu.cov()


def test_bakeoff_readme_shape(oracle):
    """cfg 1a (README.md:51-61) scaled to 5,000 intervals: A-vs-A and A-vs-B, find == count, u64 offsets."""
    from rust_lapper_b200 import synth

    (sa, ea), (sb, eb) = synth.bakeoff_sets(1, 2, 5000, 100_000)
    gpu, cpu = _both(oracle, sa, ea)
    for qs, qe in ((sa, ea), (sb, eb)):
        _assert_same_queries(gpu, cpu, qs, qe)
        off, _ = gpu.find_batch(qs, qe)
        assert int(off[-1]) == int(gpu.count_batch(qs, qe).sum())


@pytest.mark.parametrize("seed", range(3))
def test_merge_cov_union(oracle, seed):
    rng = np.random.default_rng(300 + seed)
    for _ in range(10):
        n, m = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
        U = int(rng.choice([500, 100_000]))
        s, e, qs, qe = rs.random_case(rng, n, U, int(rng.choice([0, 4, 60])), 500, 30)
        s2, e2, _, _ = rs.random_case(rng, m, U, int(rng.choice([0, 4, 60])), 1, 1)
        A, cA = _both(oracle, s, e)
        B, cB = _both(oracle, s2, e2)
        assert A.cov() == cA.cov()
        assert A.union_and_intersect(B) == cA.union_and_intersect(cB)
        assert B.union_and_intersect(A) == cB.union_and_intersect(cA)
        A.merge_overlaps()
        cA.merge_overlaps()
        assert A.overlaps_merged and len(A) == len(cA) and A.max_len == cA.max_len
        gs, ge, gp = A.intervals_soa()
        cs, ce, cv = cA.intervals_soa()
        assert np.array_equal(gs, cs) and np.array_equal(ge, ce) and np.array_equal(gp, cv)
        assert np.array_equal(A.starts, cA.starts) and np.array_equal(A.stops, cA.stops)
        assert A.cov() == cA.cov()
        _assert_same_queries(A, cA, qs, qe)
        assert A.union_and_intersect(B) == cA.union_and_intersect(cB)  # mixed flags
        B.merge_overlaps()
        cB.merge_overlaps()
        B.set_cov()
        cB.set_cov()
        assert A.union_and_intersect(B) == cA.union_and_intersect(cB)  # both merged
        A.merge_overlaps()  # idempotent
        assert np.array_equal(A.intervals_soa()[0], cs)


def test_merge_large_and_precondition(oracle):
    from rust_lapper_b200 import Lapper, LapperError, _ffi, synth

    s, e = synth.intervals_uniform(3, 300_000, 20_000_000, 0, 120)
    A, cA = _both(oracle, s, e)
    assert A.cov() == cA.cov()
    A.merge_overlaps()
    cA.merge_overlaps()
    assert np.array_equal(A.intervals_soa()[0], cA.intervals_soa()[0])
    assert np.array_equal(A.intervals_soa()[1], cA.intervals_soa()[1])
    assert np.array_equal(A.intervals_soa()[2], cA.intervals_soa()[2])
    # sets with stop < start: merge_overlaps follows the reference's loop (comparisons only, src/lib.rs:361-416);
    # cov still refuses (the reference's arithmetic overflows there)
    bad = Lapper([(10, 5), (1, 2)])
    with pytest.raises(LapperError) as ei:
        bad.cov()
    assert ei.value.code == _ffi.ERR_PRECONDITION
    bad.merge_overlaps()
    cbad = oracle.OracleLapper([(10, 5), (1, 2)])
    cbad.merge_overlaps()
    assert bad.intervals == cbad.intervals and bad.overlaps_merged


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_merge_overlaps_with_inverted_intervals(oracle, seed):
    """The reference's stack loop on sets where some stop < start: a run that opens with such an interval restarts its
    running stop BELOW stops seen earlier, so the prefix-max form does not apply; the engine runs the loop literally."""
    rng = np.random.default_rng(seed)
    n = [500, 30_000, 200_000][seed - 1]
    s = rng.integers(0, 20 * n, n).astype(np.uint32)
    e = s + rng.integers(0, 60, n).astype(np.uint32)
    k = rng.choice(n, n // 25, replace=False)
    s[k], e[k] = e[k] + np.uint32(1 + seed), s[k]
    A, cA = _both(oracle, s, e)
    assert A.has_inverted
    qs = rng.integers(0, 20 * n, 50_000).astype(np.uint32)
    qe = qs + rng.integers(0, 100, qs.size).astype(np.uint32)
    _assert_same_queries(A, cA, qs, qe)
    A.merge_overlaps(), cA.merge_overlaps()
    gs, ge, gp = A.intervals_soa()
    cs, ce, cv = cA.intervals_soa()
    assert np.array_equal(gs, cs) and np.array_equal(ge, ce) and np.array_equal(gp, cv)
    assert np.array_equal(A.stops, cA.stops) and A.max_len == cA.max_len and A.overlaps_merged
    _assert_same_queries(A, cA, qs, qe)
    A.merge_overlaps(), cA.merge_overlaps()  # again, on the merged set (which may still hold inverted runs)
    assert np.array_equal(A.intervals_soa()[0], cA.intervals_soa()[0]) and np.array_equal(A.intervals_soa()[1], cA.intervals_soa()[1])
    _assert_same_queries(A, cA, qs, qe)


def test_seek_cursor_protocol(oracle):
    rng = np.random.default_rng(21)
    s, e, qs, qe = rs.random_case(rng, 300, 20_000, 900, 120, 200)
    gpu, cpu = _both(oracle, s, e)
    # sorted queries: seek == find; unsorted: whatever the reference does with its cursor
    for order in (np.argsort(qs, kind="stable"), np.arange(qs.size)):
        cg, cc = [0], [0]
        for j in order:
            a, b = int(qs[j]), int(qe[j])
            assert gpu.seek_idx(a, b, cg) == cpu.seek_idx(a, b, cc)
            assert cg == cc


@pytest.mark.parametrize("shift", [0, 2, 5, 9, 13])
def test_queries_around_bucket_boundaries_and_last_bucket(oracle, shift):
    """Regression: a query whose stop lies in the first bucket past the last key (the sentinel) and whose
    start lies within the stop margin below it must still see the stops of that margin.  Sweeps starts and
    lengths around several bucket boundaries, including the last one, for forced bucket widths."""
    w = 1 << shift
    rng = np.random.default_rng(shift)
    top = 50 * w + w // 3
    s = rng.integers(0, top, 400).astype(np.uint32)
    e = np.minimum(s + rng.integers(0, 3 * w + 2, 400), top).astype(np.uint32)
    e[:40] = top  # many stops at the very last coordinate
    gpu, cpu = _both(oracle, s, e, 2 | ((shift + 1) << 8))
    assert gpu.bucket_shift == shift
    qs, qe = [], []
    for boundary in (w, 7 * w, 49 * w, 50 * w, 51 * w, 52 * w):
        for a in range(max(boundary - w - 2, 0), boundary + w + 3, max(w // 16, 1)):
            for ln in (0, 1, w // 8, w // 4, w // 4 + 1, w // 2, w, 2 * w + 1):
                qs.append(a)
                qe.append(a + ln)
    _assert_same_queries(gpu, cpu, np.array(qs, np.uint32), np.array(qe, np.uint32))


def test_query_shapes_on_dense_gene_exon_set(oracle):
    """300 k gene/exon-like intervals in a 40 M universe (dense: bucket width 512, frequent overflow and child
    sectors) against random, start-sorted, mixed, long and inverted queries -- count and find, bit for bit.
    (This shape caught the sentinel-bucket margin bug.)"""
    from rust_lapper_b200 import synth

    U = 40_000_000
    s, e = synth.intervals_gene_exon(44, 300_000, U)
    gpu, cpu = _both(oracle, s, e, 2)
    nq = 300_000 + 77
    rq = synth.queries_fixed_len(43, nq, U, 150)
    sq = synth.queries_sorted_fixed_len(45, nq, U, 150)
    mixed = (np.concatenate([sq[0][:100_000], rq[0][:100_077], sq[0][200_000:300_000]]),
             np.concatenate([sq[1][:100_000], rq[1][:100_077], sq[1][200_000:300_000]]))
    long_q = (rq[0], np.minimum(rq[0].astype(np.int64) + 30_000, 0xFFFFFFFF).astype(np.uint32))
    inverted = (rq[1], rq[0])
    for name, (qs, qe) in (("random", rq), ("sorted", sq), ("mixed", mixed), ("long", long_q), ("inverted", inverted)):
        assert np.array_equal(gpu.count_batch_c32(qs, qe), cpu.count_batch(qs, qe, 4).astype(np.uint32)), name
        go, gi = gpu.find_batch(qs, qe)
        co, ci = cpu.find_batch(qs, qe, 4)
        assert np.array_equal(go, co) and np.array_equal(gi, ci), name


def test_concurrent_const_calls_from_host_threads(oracle):
    """One handle shared by several host threads (the reference's shared `&Lapper`, src/lib.rs:7): concurrent
    count_batch / find_batch / cov calls must each return the oracle's answer."""
    import threading

    rng = np.random.default_rng(9)
    s, e, _, _ = rs.random_case(rng, 200_000, 50_000_000, 3000, 1, 1)
    gpu, cpu = _both(oracle, s, e)
    want_cov = cpu.cov()
    jobs = []
    for t in range(8):
        _, _, qs, qe = rs.random_case(np.random.default_rng(100 + t), 1, 50_000_000, 3000, 40_000 + 1000 * t, 400)
        jobs.append((qs, qe, cpu.count_batch(qs, qe), cpu.find_batch(qs, qe)))
    errors = []

    def worker(qs, qe, want_c, want_f):
        try:
            for _ in range(3):
                assert np.array_equal(gpu.count_batch(qs, qe), want_c)
                off, idx = gpu.find_batch(qs, qe)
                assert np.array_equal(off, want_f[0]) and np.array_equal(idx, want_f[1])
                assert gpu.cov() == want_cov
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    threads = [threading.Thread(target=worker, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


# ------------------------------------------------------------------------------------------------ round 2
def _dev_find(lib, h, qs, qe, off, sp):
    """two-phase device find into the caller's offsets tensor (any alignment); returns (offsets, indices) on the host"""
    import ctypes as C

    import torch

    from rust_lapper_b200 import _ffi

    nq = qs.numel()
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    total = C.c_uint64(0)
    _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h, vp(qs), vp(qe), nq, vp(off), C.byref(total), sp))
    idx = torch.empty(max(int(total.value), 1), dtype=torch.int32, device=qs.device)
    _ffi.check(lib.lapper_find_batch_fill_total_u32_dev(h, vp(qs), vp(qe), nq, vp(off), int(total.value), vp(idx), sp))
    torch.cuda.synchronize()
    return off.cpu().numpy().view(np.uint64), idx[: int(total.value)].cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("skew", [0, 1, 3])
def test_find_offsets_buffer_at_any_alignment(oracle, skew):
    """lapper_find_batch_offsets_u32_dev used to store offsets through 16-byte vectors whatever the
    caller's pointer; an 8-byte-aligned d_offsets (a tensor slice at an odd index) and query arrays at odd word
    offsets must work and give the same CSR."""
    import ctypes as C

    import torch

    from rust_lapper_b200 import _ffi, synth

    dev = torch.device("cuda", 0)
    lib = _ffi.lib()
    s, e = synth.intervals_gene_exon(44, 300_000, 60_000_000)
    qs, qe = synth.queries_sorted_fixed_len(45, 50_000, 60_000_000, 150)
    cpu = oracle.OracleLapper(starts=s, stops=e)
    want_o, want_i = cpu.find_batch(qs, qe, 4)
    h = C.c_void_p(0)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ts = torch.from_numpy(s.view(np.int32)).to(dev)
    te = torch.from_numpy(e.view(np.int32)).to(dev)
    _ffi.check(lib.lapper_build_u32_dev(C.c_void_p(ts.data_ptr()), C.c_void_p(te.data_ptr()), len(s), 0, 0, sp, C.byref(h)))
    try:
        nq = len(qs)
        big_s = torch.zeros(nq + 8, dtype=torch.int32, device=dev)
        big_e = torch.zeros(nq + 8, dtype=torch.int32, device=dev)
        big_s[skew:skew + nq] = torch.from_numpy(qs.view(np.int32)).to(dev)
        big_e[skew:skew + nq] = torch.from_numpy(qe.view(np.int32)).to(dev)
        off_buf = torch.empty(nq + 1 + 4, dtype=torch.int64, device=dev)
        off = off_buf[(1 if skew else 0):][: nq + 1]  # 8-byte aligned only when skewed
        got_o, got_i = _dev_find(lib, h, big_s[skew:skew + nq], big_e[skew:skew + nq], off, sp)
        assert np.array_equal(got_o, want_o) and np.array_equal(got_i, want_i)
    finally:
        lib.lapper_free(h)


@pytest.mark.parametrize("capacity_mode", ["exact", "size_query", "too_small"])
def test_find_batch_host_pipeline(oracle, capacity_mode):
    """lapper_find_batch_u32_host: host queries -> host CSR through the chunk pipeline (several chunks of 4 M queries,
    ragged last chunk), pinned and pageable buffers, the size-query and capacity-error protocols."""
    import ctypes as C

    from rust_lapper_b200 import Lapper, _ffi, synth

    lib = _ffi.lib()
    s, e = synth.intervals_gene_exon(44, 400_000, 300_000_000)
    nq = 9_000_017  # three chunks
    qs, qe = synth.queries_sorted_fixed_len(45, nq, 300_000_000, 150)
    qs[5::100_003] = 0xFFFFFFFF  # a few inverted queries at u32::MAX
    lap = Lapper(starts=s, stops=e)
    cpu = oracle.OracleLapper(starts=s, stops=e)
    want_o, want_i = cpu.find_batch(qs, qe, 8)
    total = C.c_uint64(0)
    off = np.empty(nq + 1, np.uint64)
    p = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    if capacity_mode == "size_query":
        _ffi.check(lib.lapper_find_batch_u32_host(lap.handle, p(qs), p(qe), nq, p(off), None, 0, C.byref(total)))
        assert int(total.value) == int(want_o[-1]) and np.array_equal(off, want_o)
    elif capacity_mode == "too_small":
        idx = np.zeros(int(want_o[-1]) // 2, np.uint32)
        rc = lib.lapper_find_batch_u32_host(lap.handle, p(qs), p(qe), nq, p(off), p(idx), idx.size, C.byref(total))
        assert rc == _ffi.ERR_CAPACITY and int(total.value) == int(want_o[-1]) and np.array_equal(off, want_o)
    else:
        idx = np.empty(int(want_o[-1]), np.uint32)
        _ffi.check(lib.lapper_find_batch_u32_host(lap.handle, p(qs), p(qe), nq, p(off), p(idx), idx.size, C.byref(total)))
        assert int(total.value) == int(want_o[-1])
        assert np.array_equal(off, want_o) and np.array_equal(idx, want_i)
        # empty batch
        off0 = np.full(1, 7, np.uint64)
        _ffi.check(lib.lapper_find_batch_u32_host(lap.handle, None, None, 0, p(off0), None, 0, C.byref(total)))
        assert int(off0[0]) == 0 and int(total.value) == 0


def test_query_tables_are_rebuilt_lazily_after_mutation(oracle):
    """merge_overlaps and insert replace the sorted arrays at once and leave the query tables (sectors, packed table,
    length classes) to the first count / find; the accessors and every query path must see the new set."""
    from rust_lapper_b200 import Lapper, _ffi

    rng = np.random.default_rng(11)
    s = rng.integers(0, 5_000_000, 200_000).astype(np.uint32)
    e = (s + rng.integers(1, 30, 200_000)).astype(np.uint32)
    qs = rng.integers(0, 5_000_000, 50_000).astype(np.uint32)
    qe = (qs + rng.integers(0, 200, 50_000)).astype(np.uint32)
    lap = Lapper(starts=s, stops=e, flags=_ffi.BUILD_FORCE_BUCKETS | _ffi.BUILD_FORCE_DENSE)
    cpu = oracle.OracleLapper(starts=s, stops=e)
    for step in range(3):
        assert np.array_equal(lap.count_batch(qs, qe), cpu.count_batch(qs, qe)), step
        go, gi = lap.find_batch(qs, qe)
        co, ci = cpu.find_batch(qs, qe)
        assert np.array_equal(go, co) and np.array_equal(gi, ci), step
        assert lap.bucket_shift is not None and lap.packed_table is not None
        assert lap.max_len == cpu.max_len and len(lap) == len(cpu)
        if step == 0:
            lap.merge_overlaps()
            cpu.merge_overlaps()
            assert lap.max_len == cpu.max_len and len(lap) == len(cpu) and lap.cov() == cpu.cov()
        elif step == 1:
            for a, b in ((17, 90_000), (4_999_999, 5_000_400), (0, 1)):
                lap.insert(a, b)
                cpu.insert(a, b)
            assert lap.max_len == cpu.max_len


def test_replica_of_merged_set_issues_fresh_input_positions():
    """after merge_overlaps perm keeps original input positions (which may exceed the run count); a replica or
    an imported handle must not hand one of them to the next inserted interval."""
    from rust_lapper_b200 import Lapper, LapperGroup

    ivs = [(10 * i, 10 * i + 15, i) for i in range(50)] + [(1000, 1010, 50), (2000, 2010, 51)]
    lap = Lapper(ivs)
    lap.merge_overlaps()
    _, _, perm = lap.intervals_soa()
    assert len(lap) == 3 and perm.max() == 51
    grp = LapperGroup(lap, [0])
    import ctypes as C

    from rust_lapper_b200 import _ffi

    rep = _ffi.lib().lapper_group_replica(grp._h, 0)
    pos = C.c_uint64(0)
    _ffi.check(_ffi.lib().lapper_insert_u32(C.c_void_p(rep), 3000, 3005, C.byref(pos)))
    assert int(pos.value) == 52
    grp.close()


# ------------------------------------------------------------------------------------------------ one-call device find
def _dev_find_one_call(lib, h, qs, qe, capacity, sp, fill=0x5A5A5A5A):
    """lapper_find_batch_u32_dev into caller buffers; returns (status, total, offsets, indices-buffer) on the host"""
    import ctypes as C

    import torch

    nq = qs.numel()
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    off = torch.empty(nq + 1, dtype=torch.int64, device=qs.device)
    idx = torch.full((max(capacity, 1),), fill, dtype=torch.int32, device=qs.device)
    total = C.c_uint64(0)
    st = lib.lapper_find_batch_u32_dev(h, vp(qs), vp(qe), nq, vp(off), vp(idx) if capacity else C.c_void_p(0), capacity,
                                       C.byref(total), sp)
    torch.cuda.synchronize()
    return st, int(total.value), off.cpu().numpy().view(np.uint64), idx.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("shape", ["sorted", "random", "ragged", "heavy", "inverted_set", "tiny"])
def test_find_one_call_device_api(oracle, shape):
    """lapper_find_batch_u32_dev (count pass, scan and emit in one enqueue; the emit kernel derives the offsets from the
    counts): same CSR as the oracle for start-sorted batches (tile kernel), random-order ones (every tile falls back to
    the per-warp kernel), a ragged last tile, heavy queries (>= 256 hits: find_heavy_kernel reads the offsets the tile
    kernel wrote), sets the fused form does not cover (inverted intervals: two-phase inside the call), plus the
    capacity protocol: exact fit, room to spare, too small (LAPPER_ERR_CAPACITY, offsets complete, positions untouched)
    and the size query."""
    import ctypes as C

    import torch

    from rust_lapper_b200 import _ffi, synth

    dev = torch.device("cuda", 0)
    lib = _ffi.lib()
    U = 60_000_000
    rng = np.random.default_rng(5)
    if shape == "heavy":
        s, e = synth.intervals_uniform(42, 40_000, 2_000_000, 500, 80_000)  # ~ 800 hits per query
        qs, qe = synth.queries_sorted_fixed_len(45, 5_000, 2_000_000, 150)
    else:
        s, e = synth.intervals_gene_exon(44, 300_000, U)
        if shape == "inverted_set":
            s, e = s.copy(), e.copy()
            k = rng.choice(s.size, 1000, replace=False)
            s[k], e[k] = e[k], s[k]
        nq = {"ragged": 70_001, "tiny": 37}.get(shape, 64 * 1024)
        if shape == "random":
            qs, qe = synth.queries_fixed_len(43, nq, U, 150)
        else:
            qs, qe = synth.queries_sorted_fixed_len(45, nq, U, 150)
    cpu = oracle.OracleLapper(starts=s, stops=e)
    want_o, want_i = cpu.find_batch(qs, qe, 4)
    total = int(want_o[-1])
    h = C.c_void_p(0)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ts = torch.from_numpy(s.view(np.int32)).to(dev)
    te = torch.from_numpy(e.view(np.int32)).to(dev)
    _ffi.check(lib.lapper_build_u32_dev(C.c_void_p(ts.data_ptr()), C.c_void_p(te.data_ptr()), len(s), 0, 0, sp, C.byref(h)))
    try:
        tq = torch.from_numpy(qs.view(np.int32)).to(dev)
        tqe = torch.from_numpy(qe.view(np.int32)).to(dev)
        for capacity in (total, total + 12345):
            st, got_t, got_o, got_i = _dev_find_one_call(lib, h, tq, tqe, capacity, sp)
            assert st == _ffi.OK, lib.lapper_last_error()
            assert got_t == total and np.array_equal(got_o, want_o)
            assert np.array_equal(got_i[:total], want_i)
            assert np.all(got_i[total:] == 0x5A5A5A5A)  # nothing written past the result
        if total > 1:
            st, got_t, got_o, got_i = _dev_find_one_call(lib, h, tq, tqe, total - 1, sp)
            assert st == _ffi.ERR_CAPACITY
            assert got_t == total and np.array_equal(got_o, want_o)
            assert np.all(got_i == 0x5A5A5A5A)  # untouched
        st, got_t, got_o, _ = _dev_find_one_call(lib, h, tq, tqe, 0, sp)
        assert st == _ffi.OK  # the size query
        assert got_t == total and np.array_equal(got_o, want_o)
    finally:
        lib.lapper_free(h)


def test_find_batch_guesses_the_result_size_from_the_previous_batch(oracle):
    """find_batch on a handle sizes its position buffer from the handle's previous batch and runs one enqueue; a batch
    with far more hits per query than the previous one (guess too small) and one with far fewer must both be exact."""
    from rust_lapper_b200 import Lapper, synth

    U = 30_000_000
    s, e = synth.intervals_gene_exon(44, 200_000, U)
    gpu, cpu = Lapper(starts=s, stops=e), oracle.OracleLapper(starts=s, stops=e)
    short = synth.queries_sorted_fixed_len(45, 40_000, U, 10)
    long_ = synth.queries_sorted_fixed_len(46, 40_000, U, 200_000)
    for qs, qe in (short, long_, short, long_, long_):
        go, gi = gpu.find_batch(qs, qe)
        co, ci = cpu.find_batch(qs, qe, 4)
        assert np.array_equal(go, co) and np.array_equal(gi, ci)


@pytest.mark.parametrize("qlen", [250, 400, 640, 5000])
def test_tune_query_length_changes_the_table_not_the_answers(oracle, qlen):
    """lapper_tune_query_length: the packed table is rebuilt with a margin that covers the announced query length
    (a hint beyond 640 keeps the default margin); counts and CSR stay those of the oracle; the setting survives
    merge_overlaps and insert and is copied to replicas"""
    from rust_lapper_b200 import Lapper, LapperGroup, synth

    s, e = synth.intervals_uniform(5, 400_000, 120_000_000, 50, 5000)
    gpu = Lapper(starts=s, stops=e, flags=8)  # (the packed table is only built next to large sector tables: force it)
    cpu = oracle.OracleLapper(starts=s, stops=e)
    before = gpu.packed_table
    rng = np.random.default_rng(qlen)
    qs = rng.integers(0, 120_000_000, 300_000).astype(np.uint32)
    qe = qs + rng.integers(max(1, qlen - 30), qlen + 1, qs.size).astype(np.uint32)
    want = cpu.count_batch(qs, qe, 8)
    assert np.array_equal(gpu.count_batch(qs, qe), want)
    gpu.tune_query_length(qlen)
    after = gpu.packed_table
    assert after is not None
    if qlen <= 640:
        assert after[0] < before[0], "a wider margin leaves a narrower bucket"
    else:
        assert after == before
    assert np.array_equal(gpu.count_batch(qs, qe), want)
    o = np.argsort(qs, kind="stable")
    _assert_same_queries(gpu, cpu, qs[o][:50_000], qe[o][:50_000])
    grp = LapperGroup(gpu, [0, 0])
    assert np.array_equal(grp.count_batch(qs, qe), want)
    grp.close()
    gpu.insert(77, 99), cpu.insert(77, 99, len(s))
    assert np.array_equal(gpu.count_batch(qs, qe), cpu.count_batch(qs, qe, 8))
    if qlen <= 640:
        assert gpu.packed_table[0] < before[0], "the tables rebuilt after insert keep the announced length"
    gpu.merge_overlaps(), cpu.merge_overlaps()
    assert np.array_equal(gpu.count_batch(qs, qe), cpu.count_batch(qs, qe, 8))


@pytest.mark.parametrize("nq", [(4 << 20), (4 << 20) + 1, 13_000_003])
def test_host_count_from_pageable_arrays_staged(oracle, nq):
    """count_batch from pageable host arrays (numpy, Vec<u32>) of >= 4 Mi queries goes through the library's own pinned
    staging buffers and copy threads (api.cu count_host_staged): same counts, u64 and u32 outputs, ragged last chunk"""
    from rust_lapper_b200 import synth

    s, e = synth.intervals_uniform(9, 200_000, 80_000_000, 50, 5000)
    gpu, cpu = _both(oracle, s, e)
    rng = np.random.default_rng(nq & 0xFFFF)
    qs = rng.integers(0, 80_000_000, nq, dtype=np.uint64).astype(np.uint32)
    qe = qs + rng.integers(0, 300, nq).astype(np.uint32)
    want = cpu.count_batch(qs, qe, 8)
    assert np.array_equal(gpu.count_batch(qs, qe), want)
    assert np.array_equal(gpu.count_batch_c32(qs, qe), want.astype(np.uint32))
    assert np.array_equal(gpu.count_batch(qs, qe), want)  # again: the staging buffers come back from the cache


def test_find_batch_with_pageable_arrays_beyond_the_staging_piece(oracle):
    """find_batch from / into pageable numpy arrays of more than 64 MB each: the query upload and lapper_result_copy go
    through the library's pinned staging buffers in 32 MB pieces (api.cu copy_to_device / copy_to_host)"""
    from rust_lapper_b200 import synth

    s, e = synth.intervals_uniform(11, 100_000, 200_000_000, 500, 4000)
    gpu, cpu = _both(oracle, s, e)
    rng = np.random.default_rng(12)
    nq = 17_000_001
    qs = np.sort(rng.integers(0, 200_000_000, nq, dtype=np.uint64)).astype(np.uint32)
    qe = qs + np.uint32(150)
    go, gi = gpu.find_batch(qs, qe)
    wo, wi = cpu.find_batch(qs, qe, 8)
    assert gi.nbytes > (64 << 20) and go.nbytes > (64 << 20)
    assert np.array_equal(go, wo)
    assert np.array_equal(gi, wi)
    # the one-call host pipeline with the same pageable arrays
    off = np.zeros(nq + 1, np.uint64)
    idx = np.zeros(int(wo[-1]), np.uint32)
    assert gpu.find_batch_into(qs, qe, off, idx) == int(wo[-1])
    assert np.array_equal(off, wo) and np.array_equal(idx, wi)
