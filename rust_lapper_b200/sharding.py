"""Query sharding across ranks (one process per GPU, torch.distributed).

The index is replicated; queries are independent units, so the path shards with NO data-path
collective: rank r owns the contiguous slice [r*nq/G, (r+1)*nq/G) and `count` just writes its slice.
`find` produces a per-rank CSR; one all-gather of G hit totals rebases the offsets so that the
concatenation over ranks is byte-identical to the single-GPU result.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_bounds(nq: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice of [0, nq) owned by `rank` (same arithmetic as lapper_group_* in csrc/api.cu)."""
    return nq * rank // world_size, nq * (rank + 1) // world_size


def exclusive_bases(totals: List[int]) -> List[int]:
    out, run = [], 0
    for t in totals:
        out.append(run)
        run += int(t)
    return out


def gather_totals(local_total: int, group=None) -> List[int]:
    """all-gather of one u64 per rank (the only exchange `find` needs)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [int(local_total)]
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([int(local_total)], dtype=torch.int64, device=dev)
    allt = [torch.zeros_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(allt, mine, group=group)
    return [int(t.item()) for t in allt]


def rebase_offsets(local_offsets: np.ndarray, rank: int, totals: List[int]) -> np.ndarray:
    """Global CSR offsets of this rank's queries: local offsets + hits of all earlier ranks.
    The last rank keeps its closing entry; earlier ranks drop theirs (it equals the next rank's first)."""
    base = exclusive_bases(totals)[rank]
    return np.asarray(local_offsets, dtype=np.uint64) + np.uint64(base)


def assemble_csr(per_rank: List[Tuple[np.ndarray, np.ndarray]]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate per-rank (local offsets[nq_r+1], indices) into the global CSR."""
    totals = [int(o[-1]) for o, _ in per_rank]
    offs = []
    for r, (o, _) in enumerate(per_rank):
        g = rebase_offsets(o, r, totals)
        offs.append(g if r == len(per_rank) - 1 else g[:-1])
    return np.concatenate(offs), np.concatenate([i for _, i in per_rank])
