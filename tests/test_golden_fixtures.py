"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must keep
reproducing them (CPU), and the CUDA engine must reproduce them through the C ABI (GPU) -- independent of
the oracle being importable on the box."""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
HAS_DEPTH = lambda g: g["depth"].shape[0] > 0  # noqa: E731


def _check(make, g):
    lap = make(g["starts"], g["stops"])
    s, e, p = lap.intervals_soa()
    assert np.array_equal(s, g["sorted_starts"]) and np.array_equal(e, g["sorted_stops"]) and np.array_equal(p, g["perm"])
    assert lap.max_len == int(g["max_len"])
    assert np.array_equal(lap.count_batch(g["q_start"], g["q_stop"]), g["counts"])
    off, idx = lap.find_batch(g["q_start"], g["q_stop"])
    assert np.array_equal(off, g["offsets"]) and np.array_equal(idx, g["indices"])
    assert lap.cov() == int(g["cov"])
    other = make(g["q_start"], g["q_stop"])
    assert list(lap.union_and_intersect(other)) == g["union_intersect_vs_queries"].tolist()
    if HAS_DEPTH(g):
        assert np.array_equal(np.array(lap.depth(), dtype=np.uint32).reshape(-1, 3), g["depth"])
    lap.merge_overlaps()
    s, e, p = lap.intervals_soa()
    assert np.array_equal(s, g["merged_starts"]) and np.array_equal(e, g["merged_stops"]) and np.array_equal(p, g["merged_perm"])


def test_golden_files_exist():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_reproduces_golden(oracle, path):
    _check(lambda s, e: oracle.OracleLapper(starts=s, stops=e), np.load(path))


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, 1], ids=["buckets", "plain"])
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_engine_reproduces_golden(path, flags):
    from rust_lapper_b200 import Lapper

    _check(lambda s, e: Lapper(starts=s, stops=e, flags=flags), np.load(path))


@pytest.mark.gpu
def test_size_limit_is_a_status_code():
    """n >= 2^32 - 16 cannot be indexed with u32 positions: rejected up front, nothing is read or allocated."""
    import ctypes as C

    from rust_lapper_b200 import _ffi

    lib = _ffi.lib()
    dummy = np.zeros(4, np.uint32)
    h = C.c_void_p(0)
    rc = lib.lapper_build_u32(C.c_void_p(dummy.ctypes.data), C.c_void_p(dummy.ctypes.data), (1 << 32) - 8, 0, C.byref(h))
    assert rc == _ffi.ERR_INVALID_ARG and not h.value
