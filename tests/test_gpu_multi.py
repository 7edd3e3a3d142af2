"""Multi-GPU in one process (lapper_replicate / lapper_group_*): index replicated by peer copies,
queries sharded contiguously, results byte-identical to the single-GPU call.  Skipped with one GPU."""
import ctypes as C

import numpy as np
import pytest

import reference_suite as rs

pytestmark = pytest.mark.gpu


def _ngpus():
    from rust_lapper_b200 import _ffi

    c = C.c_int(0)
    _ffi.lib().lapper_device_count(C.byref(c))
    return c.value


def test_group_single_device_matches(oracle):
    """The group path with G = 1 (always runnable) and, when present, G = all devices."""
    from rust_lapper_b200 import Lapper, LapperGroup, synth

    s, e = synth.intervals_gene_exon(44, 50_000, 20_000_000)
    qs, qe = synth.queries_fixed_len(43, 70_001, 20_000_000, 150)
    lap = Lapper(starts=s, stops=e)
    cpu = oracle.OracleLapper(starts=s, stops=e)
    want_c = cpu.count_batch(qs, qe)
    want_o, want_i = cpu.find_batch(qs, qe)
    n = _ngpus()
    for devices in ([0], [0, 0, 0], list(range(n))):
        grp = LapperGroup(lap, devices)
        assert len(grp) == len(devices)
        assert np.array_equal(grp.count_batch(qs, qe), want_c)
        off, idx = grp.find_batch(qs, qe)
        assert np.array_equal(off, want_o) and np.array_equal(idx, want_i)
        grp.close()


def test_group_after_merge(oracle):
    from rust_lapper_b200 import Lapper, LapperGroup

    rng = np.random.default_rng(3)
    s, e, qs, qe = rs.random_case(rng, 3000, 100_000, 80, 5000, 40)
    lap = Lapper(starts=s, stops=e)
    cpu = oracle.OracleLapper(starts=s, stops=e)
    lap.merge_overlaps()
    cpu.merge_overlaps()
    grp = LapperGroup(lap, list(range(max(_ngpus(), 1))))
    assert np.array_equal(grp.count_batch(qs, qe), cpu.count_batch(qs, qe))
    off, idx = grp.find_batch(qs, qe)
    co, ci = cpu.find_batch(qs, qe)
    assert np.array_equal(off, co) and np.array_equal(idx, ci)


def test_group_set_operations(oracle):
    """cov summed over the replicas' chunks (one chunk per replica, seams inside runs and between them), the cached value,
    merge_overlaps on every replica: equal to the oracle on G = 1, G = 3 replicas of one device and G = all devices"""
    from rust_lapper_b200 import Lapper, LapperGroup, synth

    s, e = synth.intervals_gene_exon(44, 60_001, 3_000_000)  # dense: long runs that cross the chunk seams
    qs, qe = synth.queries_sorted_fixed_len(45, 20_000, 3_000_000, 150)
    lap = Lapper(starts=s, stops=e)
    n = _ngpus()
    for devices in ([0], [0, 0, 0], list(range(n)), list(range(n)) * 2 + [0]):
        cpu = oracle.OracleLapper(starts=s, stops=e)
        grp = LapperGroup(lap, devices)
        assert grp.cov() == cpu.cov() == lap.cov()
        assert grp.set_cov() == cpu.set_cov()
        assert grp.cov() == cpu.cov()  # the cached value (src/lib.rs:311-316)
        grp.merge_overlaps()
        cpu.merge_overlaps()
        cs, ce, cv = cpu.intervals_soa()
        for g in range(len(devices)):
            gs, ge, gp = grp.replica_intervals(g)
            assert np.array_equal(gs, cs) and np.array_equal(ge, ce) and np.array_equal(gp, cv)
        assert grp.cov() == cpu.cov()
        assert np.array_equal(grp.count_batch(qs, qe), cpu.count_batch(qs, qe))
        off, idx = grp.find_batch(qs, qe)
        co, ci = cpu.find_batch(qs, qe)
        assert np.array_equal(off, co) and np.array_equal(idx, ci)
        grp.close()
    assert not lap.overlaps_merged  # the source handle is untouched
