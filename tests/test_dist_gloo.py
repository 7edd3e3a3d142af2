"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous query shards, no collective for
count, one all-gather of hit totals for find, concatenation byte-identical to the unsharded result.
The per-shard compute here is the oracle (test infrastructure); on the GPU box the same logic runs
over the CUDA engine (bench.py --gpus N, tests/test_gpu_multi.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import lapper_oracle
    from rust_lapper_b200 import sharding, synth

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        s, e = synth.intervals_gene_exon(44, 20_000, 10_000_000)
        nq = 30_001  # not divisible by the world size
        lo, hi = sharding.shard_bounds(nq, world, rank)
        # each rank generates only its own slice of the query stream
        qs, qe = synth.queries_sorted_fixed_len(45, nq, 10_000_000, 150, offset=lo, n=hi - lo)
        lap = lapper_oracle.OracleLapper(starts=s, stops=e)
        counts = lap.count_batch(qs, qe)
        off, idx = lap.find_batch(qs, qe)
        totals = sharding.gather_totals(int(off[-1]))
        assert len(totals) == world and totals[rank] == int(off[-1])
        goff = sharding.rebase_offsets(off, rank, totals)
        np.savez(os.path.join(tmpdir, f"rank{rank}.npz"), counts=counts, off=goff, idx=idx, lo=lo, hi=hi)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded(tmp_path, world, oracle):
    from rust_lapper_b200 import sharding, synth

    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    s, e = synth.intervals_gene_exon(44, 20_000, 10_000_000)
    nq = 30_001
    qs, qe = synth.queries_sorted_fixed_len(45, nq, 10_000_000, 150)
    lap = oracle.OracleLapper(starts=s, stops=e)
    want_counts = lap.count_batch(qs, qe)
    want_off, want_idx = lap.find_batch(qs, qe)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert [int(p["lo"]) for p in parts] == [sharding.shard_bounds(nq, world, r)[0] for r in range(world)]
    assert int(parts[-1]["hi"]) == nq
    counts = np.concatenate([p["counts"] for p in parts])
    off = np.concatenate([p["off"][:-1] if r < world - 1 else p["off"] for r, p in enumerate(parts)])
    idx = np.concatenate([p["idx"] for p in parts])
    assert np.array_equal(counts, want_counts)
    assert np.array_equal(off, want_off)
    assert np.array_equal(idx, want_idx)


def test_assemble_csr_matches(oracle):
    from rust_lapper_b200 import sharding

    rng = np.random.default_rng(0)
    s = rng.integers(0, 1000, 200).astype(np.uint32)
    e = (s + rng.integers(0, 50, 200)).astype(np.uint32)
    qs = rng.integers(0, 1000, 101).astype(np.uint32)
    qe = (qs + 20).astype(np.uint32)
    lap = oracle.OracleLapper(starts=s, stops=e)
    want = lap.find_batch(qs, qe)
    for G in (1, 2, 4, 8):
        parts = []
        for r in range(G):
            lo, hi = sharding.shard_bounds(qs.size, G, r)
            parts.append(lap.find_batch(qs[lo:hi], qe[lo:hi]))
        off, idx = sharding.assemble_csr(parts)
        assert np.array_equal(off, want[0]) and np.array_equal(idx, want[1])
