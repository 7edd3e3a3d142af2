"""Seeded synthetic intervals / queries for the BASELINE.json configs.

A counter-based generator (splitmix64 of seed/stream/index), integer-only, so that numpy on the host
(oracle side) and torch on the device (GPU side) produce bit-identical arrays for any slice
[offset, offset + n) without materialising the rest -- a rank generates just its query shard.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

U32_MAX = 0xFFFFFFFF
GRCH38 = 3_100_000_000  # "3.1 Gbp GRCh38-sized universe"

_C0 = 0x9E3779B97F4A7C15
_C1 = 0xBF58476D1CE4E5B9
_C2 = 0x94D049BB133111EB


def _s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


# ------------------------------------------------------------------ numpy backend
def _mix_np(ctr: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = ctr + np.uint64(_C0)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_C1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_C2)
        return z ^ (z >> np.uint64(31))


def rand_u32_np(seed: int, stream: int, offset: int, n: int) -> np.ndarray:
    """uniform u32 words: word i = hi32(splitmix64(seed * 2^40-ish key + stream key + i))"""
    with np.errstate(over="ignore"):
        base = np.uint64((seed * 0xD1342543DE82EF95 + stream * 0xA0761D6478BD642F) & ((1 << 64) - 1))
        ctr = np.arange(offset, offset + n, dtype=np.uint64) + base
    return (_mix_np(ctr) >> np.uint64(32)).astype(np.uint64)


def uniform_np(seed: int, stream: int, offset: int, n: int, lo: int, hi: int) -> np.ndarray:
    """integers in [lo, hi) (hi - lo < 2^32) by multiply-shift of a uniform u32 word"""
    w = rand_u32_np(seed, stream, offset, n)
    with np.errstate(over="ignore"):
        return (np.uint64(lo) + ((w * np.uint64(hi - lo)) >> np.uint64(32))).astype(np.uint64)


# ------------------------------------------------------------------ torch backend (any device)
def _lsr(z, k: int):
    return (z >> k) & ((1 << (64 - k)) - 1)


def _mix_t(ctr):
    z = ctr + _s64(_C0)
    z = (z ^ _lsr(z, 30)) * _s64(_C1)
    z = (z ^ _lsr(z, 27)) * _s64(_C2)
    return z ^ _lsr(z, 31)


def rand_u32_t(seed: int, stream: int, offset: int, n: int, device):
    import torch

    base = _s64(seed * 0xD1342543DE82EF95 + stream * 0xA0761D6478BD642F)
    ctr = torch.arange(offset, offset + n, dtype=torch.int64, device=device) + base
    return _lsr(_mix_t(ctr), 32)


def uniform_t(seed: int, stream: int, offset: int, n: int, lo: int, hi: int, device):
    w = rand_u32_t(seed, stream, offset, n, device)
    return lo + _lsr(w * (hi - lo), 32)


# ------------------------------------------------------------------ distributions (backend-agnostic)
class _NP:
    @staticmethod
    def uniform(seed, stream, offset, n, lo, hi):
        return uniform_np(seed, stream, offset, n, lo, hi).astype(np.int64)

    @staticmethod
    def where(c, a, b):
        return np.where(c, a, b)

    @staticmethod
    def to_u32(x):
        return x.astype(np.uint32)


class _T:
    def __init__(self, device):
        self.device = device

    def uniform(self, seed, stream, offset, n, lo, hi):
        return uniform_t(seed, stream, offset, n, lo, hi, self.device)

    @staticmethod
    def where(c, a, b):
        import torch

        return torch.where(c, a, b)

    @staticmethod
    def to_u32(x):
        import torch

        # int32 storage holds the same 32 bits (the int64 -> int32 cast truncates); the C ABI only sees bytes
        return x.to(torch.int32)


def _backend(device):
    return _NP() if device is None else _T(device)


def intervals_uniform(seed: int, n: int, universe: int, len_lo: int, len_hi: int, device=None, offset: int = 0):
    """cfg 2 / cfg 5 / cfg 1: start ~ U[0, universe - len_hi), len ~ U[len_lo, len_hi] (inclusive)."""
    B = _backend(device)
    s = B.uniform(seed, 0, offset, n, 0, universe - len_hi)
    ln = B.uniform(seed, 1, offset, n, len_lo, len_hi + 1)
    return B.to_u32(s), B.to_u32(s + ln)


def intervals_gene_exon(seed: int, n: int, universe: int = GRCH38, device=None, offset: int = 0):
    """cfg 3: 90 % exon-like len ~ U[50, 500]; 10 % gene-like, piecewise log-uniform over [1e3, 1.28e5):
    len = 1000*2^o + U[0, 1000*2^o), o ~ U{0..6}.  Integer-only so host and device agree bit for bit."""
    B = _backend(device)
    kind = B.uniform(seed, 2, offset, n, 0, 10)
    exon = B.uniform(seed, 1, offset, n, 50, 501)
    octave = B.uniform(seed, 3, offset, n, 0, 7)
    frac = B.uniform(seed, 4, offset, n, 0, 1 << 20)
    base = 1000 * (1 << octave) if device is None else 1000 * (2 ** octave)
    gene = base + ((base * frac) >> 20)
    ln = B.where(kind == 0, gene, exon)
    s = B.uniform(seed, 0, offset, n, 0, universe - 128_000)
    return B.to_u32(s), B.to_u32(s + ln)


def intervals_pathological(seed: int, n: int, universe: int = GRCH38, device=None, offset: int = 0):
    """cfg 4: cfg-2 intervals with 1 % replaced by len ~ U[0.5 U, 0.9 U] (start chosen so stop <= U)."""
    B = _backend(device)
    s = B.uniform(seed, 0, offset, n, 0, universe - 5000)
    ln = B.uniform(seed, 1, offset, n, 50, 5001)
    pick = B.uniform(seed, 2, offset, n, 0, 100)
    big = B.uniform(seed, 3, offset, n, universe // 2, universe // 10 * 9)
    pos = B.uniform(seed, 4, offset, n, 0, 1 << 20)
    big_s = ((universe - big) * pos) >> 20
    is_big = pick == 0
    s = B.where(is_big, big_s, s)
    ln = B.where(is_big, big, ln)
    return B.to_u32(s), B.to_u32(s + ln)


def queries_fixed_len(seed: int, nq: int, universe: int, qlen: int, device=None, offset: int = 0):
    """random-order queries: start ~ U[0, universe - qlen), stop = start + qlen"""
    B = _backend(device)
    s = B.uniform(seed, 0, offset, nq, 0, universe - qlen)
    return B.to_u32(s), B.to_u32(s + qlen)


def queries_sorted_fixed_len(seed: int, nq_total: int, universe: int, qlen: int, device=None, offset: int = 0,
                             n: int | None = None):
    """start-sorted queries without a sort: query j starts in stratum j of a regular grid,
    start_j = floor(j * span / nq_total) + jitter_j with jitter below the stratum width, so the
    sequence is non-decreasing by construction and any slice [offset, offset+n) can be generated
    independently."""
    B = _backend(device)
    n = nq_total - offset if n is None else n
    span = universe - qlen
    step = max(span // nq_total, 1)
    if device is None:
        j = np.arange(offset, offset + n, dtype=np.int64)
    else:
        import torch

        j = torch.arange(offset, offset + n, dtype=torch.int64, device=device)
    # floor(j * span / nq_total) without overflow: j < 2^31, span < 2^32 -> product < 2^63
    base = (j * span) // nq_total
    jit = B.uniform(seed, 0, offset, n, 0, step)
    s = base + jit
    return B.to_u32(s), B.to_u32(s + qlen)


def to_numpy_u32(x) -> np.ndarray:
    if isinstance(x, np.ndarray):
        return x.astype(np.uint32, copy=False)
    import torch

    if x.dtype == torch.int32:
        return x.cpu().numpy().view(np.uint32)
    return x.cpu().numpy().astype(np.uint32, copy=False)


def bakeoff_sets(seed_a: int = 1, seed_b: int = 2, n: int = 50_000, universe: int = 100_000) -> Tuple:
    """cfg 1a (README.md:51-61): n intervals, start ~ U[0, universe), len ~ U[500, 80000), plus one
    universe-spanning interval; two sets A and B."""
    out = []
    for seed in (seed_a, seed_b):
        s = uniform_np(seed, 0, 0, n, 0, universe).astype(np.int64)
        ln = uniform_np(seed, 1, 0, n, 500, 80000).astype(np.int64)
        s = np.concatenate([s, [0]])
        e = np.concatenate([s[:-1] + ln, [universe]])
        out.append((s.astype(np.uint32), e.astype(np.uint32)))
    return tuple(out)
