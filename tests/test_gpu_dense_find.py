"""The dense shape of the per-warp emit kernel (csrc/query.cu, FillDense: batches with >= 20 hits per query on average)
against the oracle, on each of its candidate paths: the single sorted vector (short max_len), the length classes (a few
long intervals stretch the reference's window beyond 512 entries), the per-lane merge (queries without locality, staged
and -- with more than 2048 hits under one warp -- unstaged), heavy queries mixed in, the ragged last chunk, and the same
set through the two-phase calls and the host pipeline."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

U = 60_000_000


def _dense_set(rng, n=1_000_000, long_frac=0.0, long_len=2_000_000):
    s = rng.integers(0, U, n, dtype=np.uint64)
    ln = rng.integers(50, 5000, n, dtype=np.uint64)
    if long_frac:
        k = rng.random(n) < long_frac
        ln[k] = rng.integers(long_len // 2, long_len, int(k.sum()), dtype=np.uint64)
    return s.astype(np.uint32), np.minimum(s + ln, 0xFFFFFFFE).astype(np.uint32)


def _check(gpu, cpu, qs, qe):
    wo, wi = cpu.find_batch(qs, qe, 8)
    go, gi = gpu.find_batch(qs, qe)
    assert np.array_equal(go, wo)
    assert np.array_equal(gi, wi)
    return int(wo[-1])


@pytest.mark.parametrize("case", ["single_vector", "classes", "random_order", "crowded", "mixed_heavy"])
def test_dense_shape_paths(oracle, case):
    from rust_lapper_b200 import Lapper

    rng = np.random.default_rng({"single_vector": 1, "classes": 2, "random_order": 3, "crowded": 4, "mixed_heavy": 5}[case])
    nq = 150_001  # (ragged last chunk)
    if case == "classes":
        s, e = _dense_set(rng, long_frac=0.0005)  # max_len ~ 2 M: the single vector's window is tens of thousands long
    elif case == "mixed_heavy":
        # 3000 long intervals over the first quarter of the universe: queries there have hundreds to thousands of hits
        # (find_heavy_kernel), the rest of the batch stays with the lanes
        s, e = _dense_set(rng)
        ls = rng.integers(0, U // 8, 3000, dtype=np.uint64)
        le = ls + rng.integers(U // 8, U // 4, 3000, dtype=np.uint64)
        s = np.concatenate([s, ls.astype(np.uint32)])
        e = np.concatenate([e, le.astype(np.uint32)])
    else:
        s, e = _dense_set(rng)
    gpu, cpu = Lapper(starts=s, stops=e), oracle.OracleLapper(starts=s, stops=e)
    if case == "random_order":
        qs = rng.integers(0, U, nq, dtype=np.uint64).astype(np.uint32)
        qe = qs + rng.integers(1, 400, nq).astype(np.uint32)
    elif case == "crowded":
        # 600-bp steps and 3-kbp queries: > 2048 hits under one warp (the stage does not hold them) and > 256 candidates
        qs = (np.arange(nq, dtype=np.uint64) * 600 % (U - 10_000)).astype(np.uint32)
        qs.sort()
        qe = qs + np.uint32(3000)
    else:
        qs = np.sort(rng.integers(0, U, nq, dtype=np.uint64)).astype(np.uint32)
        qe = qs + np.uint32(150)
        k = rng.choice(nq, nq // 20, replace=False)  # empty and inverted queries in between
        qe[k] = qs[k] - rng.integers(0, 2, k.size).astype(np.uint32).clip(max=qs[k])
    total = _check(gpu, cpu, qs, qe)
    assert total >= 20 * nq, "the batch is meant to take the dense shape"
    if case == "single_vector":
        # the same batch through the two-phase device calls and the host pipeline
        import ctypes as C

        from rust_lapper_b200 import _ffi

        off = np.empty(nq + 1, np.uint64)
        idx = np.empty(total, np.uint32)
        t = gpu.find_batch_into(qs, qe, off, idx)
        wo, wi = cpu.find_batch(qs, qe, 8)
        assert t == total and np.array_equal(off, wo) and np.array_equal(idx, wi)
        assert np.array_equal(gpu.count_batch(qs, qe), cpu.count_batch(qs, qe, 8))
        del C, _ffi


def test_sparse_batch_on_the_dense_set_keeps_the_sparse_shape(oracle):
    """queries mostly outside the intervals: < 20 hits per query on average, the sparse shape answers"""
    from rust_lapper_b200 import Lapper

    rng = np.random.default_rng(77)
    s, e = _dense_set(rng, n=300_000)
    gpu, cpu = Lapper(starts=s, stops=e), oracle.OracleLapper(starts=s, stops=e)
    qs = np.sort(rng.integers(0, U, 100_000, dtype=np.uint64)).astype(np.uint32)
    qe = qs + np.uint32(1)
    total = _check(gpu, cpu, qs, qe)
    assert total < 20 * qs.size
