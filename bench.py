#!/usr/bin/env python
"""bench.py -- overlap queries/sec (count, find) of the batched rust-lapper query path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

BASELINE.json's metric is "overlap queries/sec (count, find) at 1/2/4/8 B200; HBM GB/s vs peak".  One JSON line:

  headline (`value`, `roofline`, `e2e`, `cpu_baseline`)  cfg 2: BITS count, 10 M intervals x 100 M random-order
      150-bp queries PER GPU (weak scaling: index replicated, every rank owns its own slice of the query stream).
  `find`   cfg 3: find with CSR output, 10 M gene/exon-like intervals x 100 M start-sorted 150-bp queries PER GPU
      (weak; timed on the device), with its own roofline, e2e (host pointers -> host CSR through
      lapper_find_batch_u32_host) and cpu_baseline; plus `verify`: the FIXED 100 M-query cfg-3 batch sharded over
      the ranks (shard_bounds + all-gather of shard totals), whose 64-bit CSR hash is the same at every N.
  `cfg5`   cfg 5: 50 M intervals replicated x 1 B random-order queries IN ALL, sharded over the ranks (strong:
      per-rank work shrinks with N), count hash equal at every N; merge_overlaps + cov on every replica; a sharded
      find batch with its CSR hash.
  `group_selfcheck` (when the process sees > 1 GPU): lapper_replicate + lapper_group_count/find_batch over real peers
      against the single-GPU answer.
  `extras` (N = 1 only): other shapes of the same path.

A "step" is one pass of the call over one batch of synthetic queries; device-resident numbers are timed with CUDA
events on the launching stream, max over ranks; e2e numbers include the H2D of the inputs from pinned host memory and
the D2H of the results.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_INTERVALS = 10_000_000
NQ_PER_GPU = 100_000_000
UNIVERSE = 3_100_000_000
QLEN = 150
SEED_IV, SEED_Q = 42, 43
SEED_IV3, SEED_Q3 = 44, 45
METRIC = "overlap queries/sec (count_batch, BITS)"
UNIT = "queries/s"
WORKLOAD = "cfg2: BITS count, 10M intervals (len 50-5000) in 3.1 Gbp universe x 100M random-order 150-bp queries per GPU"
WORKLOAD3 = "cfg3: find with CSR output, 10M gene/exon-like intervals x 100M start-sorted 150-bp queries per GPU"
CFG5_INTERVALS = 50_000_000
CFG5_QUERIES = 1_000_000_000
CFG5_FIND_QUERIES = 20_000_000


def packed_info(lib, h):
    """bucket width / sectors / overflowed sectors of the packed count table (None when it was not built)"""
    w, ns, no = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
    lib.lapper_packed_table_info(h, C.byref(w), C.byref(ns), C.byref(no))
    return {"bucket_width": int(w.value), "sectors": int(ns.value), "overflowed": int(no.value)} if w.value else None


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU (NVML) while the timed region runs."""

    REASONS = {
        "hw_slowdown": 0x8,
        "sw_power_cap": 0x4,
        "hw_thermal_slowdown": 0x40,
        "sw_thermal_slowdown": 0x20,
        "hw_power_brake": 0x80,
    }

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _poll(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def merge(self, other):
        self.samples += other.samples
        self.reasons |= other.reasons

    def summary(self):
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ---------------------------------------------------------------------------------- helpers
def _vp(t):
    return C.c_void_p(t.data_ptr())


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_oracle_native():
    """The oracle built -march=native for this host when gcc is here, else the shipped build."""
    from oracle import lapper_oracle

    try:
        import tempfile

        out = os.path.join(tempfile.mkdtemp(prefix="lapper_oracle_"), "liblapper_oracle_native.so")
        if lapper_oracle.build_native(out):
            return lapper_oracle, lapper_oracle.load(out), "gcc -O3 -march=native"
    except Exception:
        pass
    return lapper_oracle, lapper_oracle.load(), "gcc -O3 -march=x86-64-v3"


def _sized_run(fn, n_cal, n_max, budget_s):
    """calibrate on n_cal units, then run as many units (<= n_max) as fit the time budget; returns (units, seconds, result)"""
    t0 = time.perf_counter()
    fn(n_cal)
    rate = n_cal / max(time.perf_counter() - t0, 1e-9)
    m = int(min(n_max, max(n_cal, rate * budget_s)))
    t0 = time.perf_counter()
    r = fn(m)
    return m, time.perf_counter() - t0, r


def cpu_count_baseline(n_intervals, sample_q, threads, budget_s=15.0):
    """Times the CPU restatement of rust-lapper's count loop (oracle port) on a bounded sample of the
    same seeded workload, queries sharded over `threads` std-threads sharing one immutable lapper."""
    import numpy as np

    from rust_lapper_b200 import synth

    mod, lib, how = load_oracle_native()
    s, e = synth.intervals_uniform(SEED_IV, n_intervals, UNIVERSE, 50, 5000)
    t0 = time.perf_counter()
    lap = mod.OracleLapper(starts=s, stops=e, lib=lib)
    build_s = time.perf_counter() - t0
    qs, qe = synth.queries_fixed_len(SEED_Q, sample_q, UNIVERSE, QLEN)
    m, dt, counts = _sized_run(lambda k: lap.count_batch(qs[:k], qe[:k], threads), max(sample_q // 16, 1), sample_q, budget_s)
    return {
        "value": m / dt,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "sample": f"first {m} of the {NQ_PER_GPU} random-order queries (seed {SEED_Q}) against all {n_intervals} intervals; "
                  f"C restatement of rust-lapper 1.3.0 count (not rustc output), {how}, {threads} threads; "
                  f"Lapper::new took {build_s:.1f} s single-threaded",
        "checksum": int(np.asarray(counts, dtype=np.uint64).sum()),
        "seconds": dt,
    }, (qs[:m], qe[:m], counts)


def cpu_find_baseline(n_intervals, nq_total, sample_q, threads, budget_s=12.0, parity_q=2_000_000):
    """The reference's own find loop -- `lapper.find(start, stop).count()` per query, as its bench times it
    (benches/lapper_benchmark.rs:103-117) -- on a bounded prefix of the cfg-3 query stream, all host threads;
    plus the oracle's CSR of the first `parity_q` queries for the parity check of the timed GPU output."""
    from rust_lapper_b200 import synth

    mod, lib, how = load_oracle_native()
    s, e = synth.intervals_gene_exon(SEED_IV3, n_intervals, UNIVERSE)
    t0 = time.perf_counter()
    lap = mod.OracleLapper(starts=s, stops=e, lib=lib)
    build_s = time.perf_counter() - t0
    qs, qe = synth.queries_sorted_fixed_len(SEED_Q3, nq_total, UNIVERSE, QLEN, n=sample_q)
    m, dt, hits = _sized_run(lambda k: lap.find_count_sum(qs[:k], qe[:k], threads), max(sample_q // 16, 1), sample_q, budget_s)
    p = min(parity_q, sample_q)
    off, idx = lap.find_batch(qs[:p], qe[:p], threads)
    return {
        "value": m / dt,
        "unit": UNIT,
        "hits_per_s": hits / dt,
        "cores": threads,
        "kind": "port",
        "sample": f"first {m} of the {nq_total} start-sorted queries (seed {SEED_Q3}) against all {n_intervals} gene/exon-like "
                  f"intervals; C restatement of rust-lapper 1.3.0 `find(..).count()` (the loop of benches/lapper_benchmark.rs:103-117; "
                  f"not rustc output), {how}, {threads} threads; Lapper::new took {build_s:.1f} s single-threaded",
        "hits": int(hits),
        "seconds": dt,
    }, (p, off, idx)


# ---------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    import numpy as np

    from rust_lapper_b200 import synth

    threads = os.cpu_count() or 1
    mod, lib, how = load_oracle_native()
    s, e = synth.intervals_uniform(SEED_IV, N_INTERVALS, UNIVERSE, 50, 5000)
    lap = mod.OracleLapper(starts=s, stops=e, lib=lib)
    sample = 20_000_000
    qs, qe = synth.queries_fixed_len(SEED_Q, sample, UNIVERSE, QLEN)
    # size one step to roughly 4 s of host work so K steps + W warm-ups end within a few minutes
    t0 = time.perf_counter()
    lap.count_batch(qs[: sample // 20], qe[: sample // 20], threads)
    rate = (sample // 20) / max(time.perf_counter() - t0, 1e-9)
    m = int(min(sample, max(sample // 20, rate * 4.0)))
    for _ in range(args.warmup):
        lap.count_batch(qs[:m], qe[:m], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        counts = lap.count_batch(qs[:m], qe[:m], threads)
    dt = time.perf_counter() - t0
    value = m * args.steps / dt
    del lap, qs, qe

    # the same arm for find (cfg 3): the reference's `find(..).count()` loop on a bounded prefix of the sorted stream
    find = None
    if not args.no_find:
        s3, e3 = synth.intervals_gene_exon(SEED_IV3, N_INTERVALS, UNIVERSE)
        lap3 = mod.OracleLapper(starts=s3, stops=e3, lib=lib)
        sample3 = 20_000_000
        q3s, q3e = synth.queries_sorted_fixed_len(SEED_Q3, NQ_PER_GPU, UNIVERSE, QLEN, n=sample3)
        t0 = time.perf_counter()
        lap3.find_count_sum(q3s[: sample3 // 20], q3e[: sample3 // 20], threads)
        rate3 = (sample3 // 20) / max(time.perf_counter() - t0, 1e-9)
        m3 = int(min(sample3, max(sample3 // 20, rate3 * 2.0)))
        fsteps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(fsteps):
            hits = lap3.find_count_sum(q3s[:m3], q3e[:m3], threads)
        dt3 = time.perf_counter() - t0
        v3 = m3 * fsteps / dt3
        find = {
            "metric": "overlap queries/sec (find_batch, CSR)", "value": v3, "unit": UNIT, "hits_per_s": hits * fsteps / dt3,
            "steps": fsteps, "ms_per_step": 1e3 * dt3 / fsteps, "queries_per_step": m3, "hits_per_step": int(hits),
            "config": {"workload": WORKLOAD3, "note": "find(start, stop).count() per query (benches/lapper_benchmark.rs:103-117)"},
            "cpu_baseline": {"value": v3, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"first {m3} of the start-sorted queries per step, {how}, {threads} threads"},
            "e2e": {"value": v3, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "intervals": N_INTERVALS, "queries_per_step": m,
                   "note": "CPU restatement of rust-lapper 1.3.0 (no Rust toolchain in the image); each step is a bounded "
                           "sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"first {m} of the random-order queries per step, {how}, {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "checksum": int(np.asarray(counts, dtype=np.uint64).sum()),
    }
    if find is not None:
        line["find"] = find
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------- device-side hashing
_M64 = (1 << 64) - 1


def _hash_tensor(values, first_index: int, salt: int) -> int:
    """order-free 64-bit hash of the (global index, value) pairs of one array slice: sum over elements of
    mix(mix(index + salt) ^ value) mod 2^64 -- shards add up, so the hash of a sharded result equals the hash of the
    whole (values: integer tensor on any device; first_index: global index of its first element)."""
    import torch

    from rust_lapper_b200.synth import _mix_t, _s64

    total = 0
    n = values.numel()
    step = 1 << 26
    for lo in range(0, n, step):
        v = values[lo:lo + step].to(torch.int64)
        if values.dtype == torch.int32:  # u32 payloads travel in int32 storage
            v = v & 0xFFFFFFFF
        idx = torch.arange(first_index + lo, first_index + lo + v.numel(), dtype=torch.int64, device=values.device)
        h = _mix_t(_mix_t(idx + _s64(salt)) ^ v)
        # the sum wraps mod 2^64 in two's complement
        total = (total + (int(h.sum().item()) & _M64)) & _M64
    return total


def csr_hash(offsets, first_query: int, indices, first_hit: int, offset_base: int, with_last: bool) -> int:
    """hash of a shard of a CSR result expressed in GLOBAL terms: offsets rebased by offset_base and keyed by global
    query index, positions keyed by their global place in `indices`.  The shard that holds the last query also
    contributes the closing offset."""
    off = offsets if with_last else offsets[:-1]
    import torch

    h1 = _hash_tensor(off.to(torch.int64) + offset_base, first_query, 0x6F66667365747321)
    h2 = _hash_tensor(indices, first_hit, 0x696E646963657321)
    return (h1 + h2) & _M64


# ---------------------------------------------------------------------------------- our arm
class Ctx:
    pass


def run_ours(args):
    import numpy as np
    import torch

    from rust_lapper_b200 import _ffi, sharding, synth

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; rust_lapper_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    lib = _ffi.lib()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_hash_over_ranks(h: int) -> int:
        if dist is None:
            return h & _M64
        parts = sharding.gather_totals(h - (1 << 64) if h >= (1 << 63) else h)  # as int64
        return sum(p & _M64 for p in parts) & _M64

    def all_true(ok: bool) -> bool:
        if dist is None:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    cx = Ctx()
    cx.args, cx.lib, cx.dev, cx.dist, cx.world, cx.rank, cx.local_rank = args, lib, dev, dist, world, rank, local_rank
    cx.barrier, cx.max_over_ranks, cx.sum_hash_over_ranks, cx.all_true = barrier, max_over_ranks, sum_hash_over_ranks, all_true
    cx.stream = torch.cuda.current_stream()
    cx.sp = C.c_void_p(cx.stream.cuda_stream)
    cx.peak, cx.peak_src = hbm_peak()
    cx.clocks = ClockSampler(local_rank)

    n, nq = args.intervals, args.queries
    stream, sp = cx.stream, cx.sp

    # ---- index: replicated (every rank builds it from the same seeded intervals, on its own GPU)
    iv_s, iv_e = synth.intervals_uniform(SEED_IV, n, UNIVERSE, 50, 5000, device=dev)
    torch.cuda.synchronize()
    h = C.c_void_p(0)
    t0 = time.perf_counter()
    _ffi.check(lib.lapper_build_u32_dev(_vp(iv_s), _vp(iv_e), n, local_rank, 0, sp, C.byref(h)))
    build_ms_first = 1e3 * (time.perf_counter() - t0)
    # warm: the library pool now holds the memory, the module is loaded -- what a rebuild costs in a running process
    build_ms = None
    for _ in range(3):
        h2 = C.c_void_p(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _ffi.check(lib.lapper_build_u32_dev(_vp(iv_s), _vp(iv_e), n, local_rank, 0, sp, C.byref(h2)))
        dt = 1e3 * (time.perf_counter() - t0)
        build_ms = dt if build_ms is None else min(build_ms, dt)
        lib.lapper_free(h2)
    shift = int(lib.lapper_bucket_shift(h))
    del iv_s, iv_e

    # ---- queries: rank r owns slice [r*nq, (r+1)*nq) of the seeded stream (weak scaling)
    qs, qe = synth.queries_fixed_len(SEED_Q, nq, UNIVERSE, QLEN, device=dev, offset=rank * nq)
    counts = torch.empty(nq, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def step():
        _ffi.check(lib.lapper_count_batch_u32_dev(h, _vp(qs), _vp(qe), nq, _vp(counts), sp))

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = int(lib.lapper_query_kernel_launches())
    with cx.clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        torch.cuda.synchronize()
    launches_timed = int(lib.lapper_query_kernel_launches()) - launches0  # counted by the library, this rank
    barrier()
    total_ms = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = total_ms / args.steps
    value = world * nq * args.steps / (total_ms * 1e-3)
    checksum = int(counts.to(torch.int64).sum().item())
    cx.headline_checksum = checksum

    # per-launch distribution (each launch separately timed) for the roofline of the one kernel
    per_launch = []
    for _ in range(min(args.steps, 10)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step()
        b.record(stream)
        torch.cuda.synchronize()
        per_launch.append(a.elapsed_time(b))
    peak, peak_src = cx.peak, cx.peak_src
    algo_bytes = nq * 12 + n * 8  # 12 B/query + 8 B/interval
    traffic = None
    try:  # dram__bytes_read + write of the dominant kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "count_kernel_traffic.json")))
        if tj.get("intervals") == n and tj.get("queries") == nq:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    kernel_ms = total_ms / args.steps  # one working kernel per step; events bracket back-to-back launches
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm",
        "kernel": "count_bucket_kernel<u32 out>",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": traffic,
        "traffic_source": "profiles/count_kernel_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" if traffic else None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": algo_bytes,
        "launch_ms_avg": kernel_ms,
        "launch_ms_isolated_median": statistics.median(per_launch),
        "note": ("random-order queries: one 32-byte sector of the 75 MB packed table per query (csrc/packed_sector.cuh); "
                 "about 64 % of those gathers hit L2, the rest cost 64 B of DRAM each (see traffic); the ceiling of one "
                 "random sector per query with the 12 B/query streams is ~0.61 ms per 100 M (profiles/, DESIGN.md section 4)")
        if (n == N_INTERVALS and nq == NQ_PER_GPU) else
        "non-default size: a sector table far beyond L2 is bound by the random 64-byte DRAM access rate; DESIGN.md section 4",
    }

    # ---- end to end: host pointers (pinned), H2D + kernel + D2H inside the timed region
    hq_s = torch.empty(nq, dtype=torch.int32, pin_memory=True)
    hq_e = torch.empty(nq, dtype=torch.int32, pin_memory=True)
    h_out = torch.empty(nq, dtype=torch.int32, pin_memory=True)
    hq_s.copy_(qs)
    hq_e.copy_(qe)
    torch.cuda.synchronize()

    def e2e_step():
        _ffi.check(lib.lapper_count_batch_u32_c32(h, _vp(hq_s), _vp(hq_e), nq, _vp(h_out)))

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_ok = bool(torch.equal(h_out, counts.cpu()))
    e2e = {
        "value": world * nq * e2e_steps / e2e_s,
        "unit": UNIT,
        "h2d_bytes_per_step": world * nq * 8,
        "d2h_bytes_per_step": world * nq * 4,
        "steps": e2e_steps,
        "ms_per_step": 1e3 * e2e_s / e2e_steps,
        "api": "lapper_count_batch_u32_c32 (host pointers, pinned; 3-slot chunked H2D/kernel/D2H pipeline)",
        "matches_device_run": e2e_ok,
    }
    # what the host side delivers when all N GPUs copy the step's bytes at once (800 MB in + 400 MB out per GPU, nothing
    # else running): measured by tools/mb_pcie.py / tools/mb_pcie_multi.py on this pool's boxes, logs under profiles/
    ceiling_ms = PCIE_MIX_MS.get(world)
    if ceiling_ms and nq == 100_000_000:
        e2e["concurrent_copy_ceiling_ms"] = ceiling_ms
        e2e["frac_of_concurrent_copy_ceiling"] = ceiling_ms / e2e["ms_per_step"]
        e2e["ceiling_source"] = PCIE_MIX_SOURCE
    del hq_s, hq_e, h_out

    # ---- CPU baseline beside it (rank 0, N = 1 only) + parity of the timed outputs on the sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu, (cq_s, cq_e, ccounts) = cpu_count_baseline(n, min(nq, 30_000_000), os.cpu_count() or 1)
            m = len(cq_s)
            got = counts[:m].cpu().numpy().view(np.uint32).astype(np.uint64)
            cpu["gpu_matches_oracle_on_sample"] = bool(np.array_equal(got, ccounts))
            del cq_s, cq_e, ccounts
        except Exception as ex:  # the baseline is a report, never a reason to lose the GPU line
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex!r}"}
    config = {
        "workload": WORKLOAD if (n == N_INTERVALS and nq == NQ_PER_GPU) else (
            f"cfg5-style: BITS count, {n} intervals (len 50-5000) in 3.1 Gbp universe x {nq} random-order 150-bp "
            f"queries per GPU ({world * nq} in all)"),
        "intervals": n,
        "queries_per_gpu": nq,
        "packed_table": packed_info(lib, h),
        "query_order": "random (as given; no pre-sort)",
        "bucket_shift": None if shift == 0xFFFFFFFF else shift,
        "index_bytes": int(lib.lapper_index_bytes(h)),
        "index_build_ms": build_ms,
        "index_build_ms_first_call": build_ms_first,
        "index_build_note": "wall clock of lapper_build_u32_dev (sort + all query tables, synchronised): best of 3 in the running "
                            "process; first_call also pays module load and the first growth of the library's memory pool",
        "parallelism": f"index replicated x{world}, queries sharded contiguously, no data-path collective",
        "l2_policy": "inputs (1.2 GB of queries + results per step) are far larger than the 126 MB L2; no flush needed",
    }
    lib.lapper_free(h)
    del qs, qe, counts
    torch.cuda.empty_cache()

    # ---- the other half of the metric and cfg 5, at every N
    find = None if args.no_find else bench_find(cx)
    torch.cuda.empty_cache()
    cfg5 = None if not args.big_intervals else bench_cfg5(cx)
    torch.cuda.empty_cache()
    selfcheck = None if args.no_selfcheck else group_selfcheck(cx)

    extras = {}
    if not args.no_extras and world == 1:
        extras = run_extras(cx, n, nq)

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u32",
            "data": "synthetic",
            "config": config,
            "roofline": roofline,
            "e2e": e2e,
            "gpu_launches": world * launches_timed,
            "gpu_launches_note": "kernels launched by the library inside the headline's timed region, counted by "
                                 "lapper_query_kernel_launches(); per step: count_bucket_kernel at 3 CTAs/SM (does the work for "
                                 "order-less batches like this one), its 4-CTAs/SM build and count_sorted_kernel (for local / sorted "
                                 "dense batches; both read the same on-device batch probe and exit at once here), "
                                 "count_bucket_tail_kernel for the ragged tail",
            "clocks": cx.clocks.summary(),
            "checksum": checksum,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if find is not None:
            line["find"] = find
        if cfg5 is not None:
            line["cfg5"] = cfg5
        if selfcheck is not None:
            line["group_selfcheck"] = selfcheck
        if extras:
            line["extras"] = extras
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def timed_events(cx, fn, steps, warm=3):
    import torch

    for _ in range(warm):
        fn()
    cx.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(cx.stream)
    for _ in range(steps):
        fn()
    b.record(cx.stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    cx.barrier()
    return ms


# ms for one count e2e step's copies (800 MB H2D + 400 MB D2H per GPU, all GPUs at once, pinned memory), by GPU count
PCIE_MIX_MS = {1: 16.1, 2: 19.19, 4: 32.24, 8: 71.09}
PCIE_MIX_SOURCE = ("profiles/r1_mb_pcie.log (N=1), profiles/r2_mb_pcie_{2,4,8}gpu.log: 'pinned, H2D + D2H' (every GPU starts both "
                   "copies of the whole step at once; the library's staggered chunk pipeline can come in slightly under it); the host "
                   "side does not scale with N (one NUMA node, shared root complex: ~100 GB/s in + ~50 GB/s out for the whole box "
                   "from N = 4 on), so neither can e2e")


def find_device(cx, h, qs, qe, nq, offsets, indices=None):
    """two-phase device find through the C ABI; returns (total, indices tensor)"""
    import torch

    from rust_lapper_b200 import _ffi

    total = C.c_uint64(0)
    _ffi.check(cx.lib.lapper_find_batch_offsets_u32_dev(h, _vp(qs), _vp(qe), nq, _vp(offsets), C.byref(total), cx.sp))
    t = int(total.value)
    if indices is None or indices.numel() < max(t, 1):
        indices = torch.empty(max(t, 1), dtype=torch.int32, device=cx.dev)
    _ffi.check(cx.lib.lapper_find_batch_fill_total_u32_dev(h, _vp(qs), _vp(qe), nq, _vp(offsets), t, _vp(indices), cx.sp))
    return t, indices


def sharded_find_hash(cx, h, gen_slice, nq_total):
    """The fixed batch of nq_total queries, sharded contiguously over the ranks: per-rank CSR, shard totals
    all-gathered (the one exchange find needs), global 64-bit hash of (offsets, positions).  Same value at every N."""
    import torch

    from rust_lapper_b200 import sharding

    lo, hi = sharding.shard_bounds(nq_total, cx.world, cx.rank)
    m = hi - lo
    qs, qe = gen_slice(lo, m)
    offsets = torch.empty(m + 1, dtype=torch.int64, device=cx.dev)
    total, indices = find_device(cx, h, qs, qe, m, offsets)
    torch.cuda.synchronize()
    totals = sharding.gather_totals(total)
    base = sharding.exclusive_bases(totals)[cx.rank]
    hsh = csr_hash(offsets, lo, indices[:total], base, base, with_last=(cx.rank == cx.world - 1))
    return cx.sum_hash_over_ranks(hsh), sum(totals), (qs, qe, offsets, indices, total, lo)


def bench_find(cx):
    """cfg 3, the `find` half of the metric."""
    import numpy as np
    import torch

    from rust_lapper_b200 import _ffi, synth

    args, lib, dev, world, rank = cx.args, cx.lib, cx.dev, cx.world, cx.rank
    n = args.intervals
    nq = min(args.queries, args.find_queries)
    s3, e3 = synth.intervals_gene_exon(SEED_IV3, n, UNIVERSE, device=dev)
    h3 = C.c_void_p(0)
    _ffi.check(lib.lapper_build_u32_dev(_vp(s3), _vp(e3), n, cx.local_rank, 0, cx.sp, C.byref(h3)))
    del s3, e3

    # ---- verification pass: the fixed nq-query batch sharded over the ranks; hash independent of N
    gen = lambda lo, m: synth.queries_sorted_fixed_len(SEED_Q3, nq, UNIVERSE, QLEN, device=dev, offset=lo, n=m)
    vhash, vtotal, (vqs, vqe, voff, vidx, vt, vlo) = sharded_find_hash(cx, h3, gen, nq)
    verify = {"queries": nq, "hits": int(vtotal), "csr_hash": f"{vhash:016x}",
              "note": "fixed batch (seed 45) sharded contiguously over the ranks, per-rank CSR rebased by the all-gathered shard "
                      "totals; hash = sum of mix(global index, value) over global offsets and positions: equal at every N"}
    # parity of this rank's shard head against the CPU oracle (N = 1: also the CPU baseline)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu, (p, coff, cidx) = cpu_find_baseline(n, nq, min(nq, 30_000_000), os.cpu_count() or 1)
            goff = voff[: p + 1].cpu().numpy().view(np.uint64)
            gidx = vidx[: int(coff[p])].cpu().numpy().view(np.uint32)
            cpu["gpu_matches_oracle_on_sample"] = bool(np.array_equal(goff, coff) and np.array_equal(gidx, cidx))
            cpu["parity_sample"] = f"CSR (offsets + positions, byte-equal) of the first {p} queries, {int(coff[p])} hits"
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex!r}"}
    del vqs, vqe, voff, vidx

    # ---- timed, weak: every rank its own nq start-sorted queries over the whole universe (rank 0's are the fixed batch)
    qs, qe = synth.queries_sorted_fixed_len(SEED_Q3 + 1000 * rank, nq, UNIVERSE, QLEN, device=dev)
    offsets = torch.empty(nq + 1, dtype=torch.int64, device=dev)
    hits, indices = find_device(cx, h3, qs, qe, nq, offsets)
    state = {}

    def offsets_step():
        t = C.c_uint64(0)
        _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h3, _vp(qs), _vp(qe), nq, _vp(offsets), C.byref(t), cx.sp))
        state["t"] = int(t.value)

    def fill_step():
        _ffi.check(lib.lapper_find_batch_fill_total_u32_dev(h3, _vp(qs), _vp(qe), nq, _vp(offsets), state["t"], _vp(indices), cx.sp))

    def two_phase_step():
        offsets_step()
        fill_step()

    # the timed call: lapper_find_batch_u32_dev, both phases in ONE enqueue into buffers of known size (what a
    # steady-state pipeline re-uses): the emit kernel derives the offsets from the counts, one host sync at the end
    offsets1 = torch.empty_like(offsets)
    indices1 = torch.empty_like(indices)

    def find_step():
        t = C.c_uint64(0)
        _ffi.check(lib.lapper_find_batch_u32_dev(h3, _vp(qs), _vp(qe), nq, _vp(offsets1), _vp(indices1), indices1.numel(),
                                                 C.byref(t), cx.sp))
        state["t1"] = int(t.value)

    find_step()
    torch.cuda.synchronize()
    one_call_ok = state["t1"] == hits and bool(torch.equal(offsets1, offsets)) and bool(torch.equal(indices1[:hits], indices[:hits]))
    steps = max(1, min(args.steps, 10))
    launches0 = int(lib.lapper_query_kernel_launches())
    clocks = ClockSampler(cx.local_rank)
    with clocks:
        ms = cx.max_over_ranks(timed_events(cx, find_step, steps))
    cx.clocks.merge(clocks)
    launches = (int(lib.lapper_query_kernel_launches()) - launches0) // (steps + 3)
    ms_two = cx.max_over_ranks(timed_events(cx, two_phase_step, steps, warm=1))
    ms_off = cx.max_over_ranks(timed_events(cx, offsets_step, steps, warm=1))
    ms_fill = cx.max_over_ranks(timed_events(cx, fill_step, steps, warm=1))
    hits_all = hits
    if cx.dist is not None:
        t = torch.tensor([hits], dtype=torch.int64, device=dev)
        cx.dist.all_reduce(t)
        hits_all = int(t.item())
    fbytes = nq * 16 + hits * 4 + n * 8  # 8 B query in + 8 B offset out per query, 4 B per hit, index once
    gbs = fbytes / (ms * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "find_kernel_traffic.json")))
        if tj.get("intervals") == n and tj.get("queries") == nq:
            traffic = tj["dram_bytes_per_step"]
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": "find pipeline: count_sorted_kernel<find> (counts + tile sums), tile-sum scans, find_fill_tile_kernel<from counts> (offsets + emit)",
        "achieved": gbs, "peak": cx.peak, "unit": "GB/s", "frac": gbs / cx.peak, "traffic": traffic,
        "traffic_source": "profiles/find_kernel_traffic.json (ncu --set full, sum over the step's kernels)" if traffic else None,
        "peak_source": cx.peak_src, "algorithmic_bytes_per_step": fbytes, "step_ms": ms,
        "two_phase_ms": ms_two, "offsets_phase_ms": ms_off, "fill_phase_ms": ms_fill,
        "note": "this rank's figures (every rank runs the same shape); step_ms = lapper_find_batch_u32_dev (one enqueue); the "
                "two-phase calls (lapper_find_batch_offsets_u32_dev + ..._fill_total_u32_dev) timed beside it, together and "
                "separately, each incl. its host sync",
    }

    # ---- end to end: host queries (pinned) -> host CSR (pinned) through ONE call, H2D / count / emit / D2H pipelined
    e2e = None
    if not args.no_e2e_find:
        hq_s = torch.empty(nq, dtype=torch.int32, pin_memory=True)
        hq_e = torch.empty(nq, dtype=torch.int32, pin_memory=True)
        h_off = torch.empty(nq + 1, dtype=torch.int64, pin_memory=True)
        h_idx = torch.empty(max(hits, 1), dtype=torch.int32, pin_memory=True)
        hq_s.copy_(qs)
        hq_e.copy_(qe)
        torch.cuda.synchronize()
        tot = C.c_uint64(0)

        def e2e_step():
            _ffi.check(lib.lapper_find_batch_u32_host(h3, _vp(hq_s), _vp(hq_e), nq, _vp(h_off), _vp(h_idx), h_idx.numel(), C.byref(tot)))

        for _ in range(2):  # (the first call grows the library's memory pool, the second settles its chunk buffers)
            e2e_step()
        cx.barrier()
        e2e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = cx.max_over_ranks(time.perf_counter() - t0)
        cx.barrier()
        ok = int(tot.value) == hits and bool(torch.equal(h_off, offsets.cpu())) and bool(torch.equal(h_idx[:hits], indices[:hits].cpu()))
        d2h = (nq + 1) * 8 + hits * 4
        # what the link delivers for the same bytes on THIS box, now: the step's D2H bytes in 192 MB pieces with the
        # step's H2D beside it, plain copies on two streams (all ranks at once, max over ranks)
        link_ms = None
        if hits:
            s_out, s_in = torch.cuda.Stream(), torch.cuda.Stream()
            piece = 48_000_000

            def link_step():
                with torch.cuda.stream(s_out):
                    for lo in range(0, hits, piece):
                        h_idx[lo:lo + piece].copy_(indices[lo:lo + piece], non_blocking=True)
                    h_off.copy_(offsets, non_blocking=True)
                with torch.cuda.stream(s_in):
                    qs.copy_(hq_s, non_blocking=True)
                    qe.copy_(hq_e, non_blocking=True)

            torch.cuda.synchronize()
            link_step()
            torch.cuda.synchronize()
            cx.barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                link_step()
            torch.cuda.synchronize()
            link_ms = 1e3 * cx.max_over_ranks(time.perf_counter() - t0) / 2
            cx.barrier()
        e2e = {
            "value": world * nq * e2e_steps / e2e_s, "unit": UNIT, "hits_per_s": hits_all * e2e_steps / e2e_s,
            "h2d_bytes_per_step": world * nq * 8, "d2h_bytes_per_step": world * d2h if world == 1 else None,
            "d2h_bytes_per_step_this_rank": d2h,
            "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
            "api": "lapper_find_batch_u32_host (host pointers, pinned; 3-slot chunk pipeline: H2D queries | offsets + emit | D2H offsets + positions)",
            "matches_device_run": cx.all_true(ok),
            "pcie_floor_ms_at_55GBps": 1e3 * d2h / 55e9,
            "link_ms_same_bytes": link_ms,
            "frac_of_link": (link_ms / (1e3 * e2e_s / e2e_steps)) if link_ms else None,
            "link_note": "link_ms_same_bytes = plain cudaMemcpyAsync of the step's D2H bytes (192 MB pieces) with its H2D bytes beside "
                         "them, timed in this process right after the e2e steps (max over ranks, all ranks copying at once)",
        }
        if world > 1:
            t = torch.tensor([d2h], dtype=torch.int64, device=dev)
            cx.dist.all_reduce(t)
            e2e["d2h_bytes_per_step"] = int(t.item())
        del hq_s, hq_e, h_off, h_idx
    out = {
        "metric": "overlap queries/sec (find_batch, CSR offsets + positions in the reference's iteration order)",
        "value": world * nq / (ms * 1e-3), "unit": UNIT, "hits_per_s": hits_all / (ms * 1e-3),
        "n_gpus": world, "steps": steps, "ms_per_step": ms, "scaling": "weak", "higher_is_better": True,
        "config": {"workload": WORKLOAD3, "intervals": n, "queries_per_gpu": nq, "hits_per_query": hits / nq,
                   "max_len": int(lib.lapper_max_len(h3)), "query_order": "start-sorted (cfg 3 as stated)"},
        "hits_this_rank": hits, "hits": hits_all,
        "api": "lapper_find_batch_u32_dev (device buffers of known capacity; count pass + scan + emit in one enqueue)",
        "one_call_matches_two_phase": cx.all_true(one_call_ok),
        "roofline": roofline, "gpu_launches_per_step": launches,
        "verify": verify,
    }
    if e2e is not None:
        out["e2e"] = e2e
    if cpu is not None:
        out["cpu_baseline"] = cpu
    lib.lapper_free(h3)
    return out


def bench_cfg5(cx):
    """cfg 5: 50 M intervals replicated, 1 B queries in all sharded over the ranks, incl. merge_overlaps + cov."""
    import torch

    from rust_lapper_b200 import _ffi, sharding, synth

    args, lib, dev, world, rank = cx.args, cx.lib, cx.dev, cx.world, cx.rank
    n5, nq_all = args.big_intervals, args.big_queries
    lo, hi = sharding.shard_bounds(nq_all, world, rank)
    m = hi - lo
    s5, e5 = synth.intervals_uniform(SEED_IV, n5, UNIVERSE, 50, 5000, device=dev)
    torch.cuda.synchronize()

    def wall(fn, reps=3):
        best, r = None, None
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return cx.max_over_ranks(best * 1e3), r

    def build5():
        hh = C.c_void_p(0)
        _ffi.check(lib.lapper_build_u32_dev(_vp(s5), _vp(e5), n5, cx.local_rank, 0, cx.sp, C.byref(hh)))
        return hh

    # best of 3, the previous index freed first so that the library's pool hands the same memory out again (with the
    # earlier handles kept alive every build grew the pool by another 2.7 GB inside the timed region)
    best5, h5 = None, None
    for _ in range(3):
        if h5 is not None:
            lib.lapper_free(h5)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h5 = build5()
        torch.cuda.synchronize()
        dt5 = time.perf_counter() - t0
        best5 = dt5 if best5 is None else min(best5, dt5)
    ms_build = cx.max_over_ranks(best5 * 1e3)

    # ---- count: this rank's shard of the 1 B random-order queries, generated on the device in pieces
    qs = torch.empty(m, dtype=torch.int32, device=dev)
    qe = torch.empty(m, dtype=torch.int32, device=dev)
    piece = 1 << 26
    for p in range(0, m, piece):
        k = min(piece, m - p)
        a, b = synth.queries_fixed_len(SEED_Q, k, UNIVERSE, QLEN, device=dev, offset=lo + p)
        qs[p:p + k], qe[p:p + k] = a, b
    del a, b
    counts = torch.empty(m, dtype=torch.int32, device=dev)
    steps = max(1, min(args.steps, 5))
    clocks = ClockSampler(cx.local_rank)
    with clocks:
        ms = cx.max_over_ranks(timed_events(
            cx, lambda: _ffi.check(lib.lapper_count_batch_u32_dev(h5, _vp(qs), _vp(qe), m, _vp(counts), cx.sp)), steps, warm=2))
    cx.clocks.merge(clocks)
    chash = cx.sum_hash_over_ranks(_hash_tensor(counts, lo, 0x636F756E74732121))
    csum = int(counts.to(torch.int64).sum().item())
    if cx.dist is not None:
        t = torch.tensor([csum], dtype=torch.int64, device=dev)
        cx.dist.all_reduce(t)
        csum = int(t.item())
    algo = m * 12 + n5 * 8
    gbs = algo / (ms * 1e-3) / 1e9
    out = {
        "workload": f"cfg5: {n5} intervals (len 50-5000, replicated) x {nq_all} random-order 150-bp queries in all, sharded "
                    f"contiguously over {world} GPU(s)",
        "scaling": "strong", "intervals": n5, "queries": nq_all, "queries_this_rank": m,
        "count": {"value": nq_all / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
                  "counts_hash": f"{chash:016x}", "counts_sum": csum,
                  "roofline": {"bound": "hbm", "achieved": gbs, "peak": cx.peak, "unit": "GB/s", "frac": gbs / cx.peak,
                               "algorithmic_bytes_per_launch": algo,
                               "note": "per rank; the 377 MB packed table is far beyond L2: bound by random 64-byte DRAM accesses"}},
        "packed_table": packed_info(lib, h5), "index_bytes": int(lib.lapper_index_bytes(h5)),
        "build_ms": ms_build, "build_intervals_per_s": n5 / (ms_build * 1e-3),
    }
    del counts

    # ---- find: a fixed start-sorted batch sharded the same way; hash equal at every N
    nqf = args.big_find_queries
    if nqf:
        gen = lambda flo, fm: synth.queries_sorted_fixed_len(SEED_Q3, nqf, UNIVERSE, QLEN, device=dev, offset=flo, n=fm)
        fhash, ftotal, (fqs, fqe, foff, fidx, ft, flo) = sharded_find_hash(cx, h5, gen, nqf)
        fm = fqs.numel()
        st = {"t": ft}

        def fstep():
            tt = C.c_uint64(0)
            _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h5, _vp(fqs), _vp(fqe), fm, _vp(foff), C.byref(tt), cx.sp))
            _ffi.check(lib.lapper_find_batch_fill_total_u32_dev(h5, _vp(fqs), _vp(fqe), fm, _vp(foff), int(tt.value), _vp(fidx), cx.sp))

        ms_f = cx.max_over_ranks(timed_events(cx, fstep, 3, warm=1))
        fb = fm * 16 + ft * 4 + n5 * 8
        out["find"] = {"queries": nqf, "queries_this_rank": fm, "hits": int(ftotal), "csr_hash": f"{fhash:016x}",
                       "value": nqf / (ms_f * 1e-3), "unit": UNIT, "hits_per_s": ftotal / (ms_f * 1e-3), "ms_per_step": ms_f,
                       "achieved_GBps_this_rank": fb / (ms_f * 1e-3) / 1e9,
                       "note": "start-sorted 150-bp queries (seed 45) over the whole universe, sharded contiguously; per-rank CSR "
                               "rebased by the all-gathered shard totals"}
        del fqs, fqe, foff, fidx
    del qs, qe

    # ---- set operations on every replica (replicas only: 0.4 GB of input is ~0.1 ms on one GPU)
    cov = C.c_uint32(0)
    ms_cov, _ = wall(lambda: _ffi.check(lib.lapper_cov(h5, C.byref(cov))))
    s5b, e5b = synth.intervals_uniform(SEED_IV + 100, n5 // 5, UNIVERSE, 50, 5000, device=dev)
    h5b = C.c_void_p(0)
    _ffi.check(lib.lapper_build_u32_dev(_vp(s5b), _vp(e5b), n5 // 5, cx.local_rank, 0, cx.sp, C.byref(h5b)))
    u, i = C.c_uint32(0), C.c_uint32(0)
    ms_ui, _ = wall(lambda: _ffi.check(lib.lapper_union_and_intersect(h5, h5b, C.byref(u), C.byref(i))))
    # merge_overlaps mutates: time it on fresh replicas
    best_merge = None
    merged_runs = None
    for rep in range(3):
        hm = h5 if rep == 2 else build5()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _ffi.check(lib.lapper_merge_overlaps(hm))
        torch.cuda.synchronize()
        dt = 1e3 * (time.perf_counter() - t0)
        best_merge = dt if best_merge is None else min(best_merge, dt)
        merged_runs = int(lib.lapper_len(hm))
        if hm is not h5:
            lib.lapper_free(hm)
    ms_merge = cx.max_over_ranks(best_merge)
    cov2 = C.c_uint32(0)
    _ffi.check(lib.lapper_cov(h5, C.byref(cov2)))
    # a query against the merged set (its tables are built lazily, on this first call)
    q1 = torch.tensor([1000, 1_000_000, 2_000_000_000], dtype=torch.int32, device=dev)
    q2 = q1 + 150
    c1 = torch.empty(3, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _ffi.check(lib.lapper_count_batch_u32_dev(h5, _vp(q1), _vp(q2), 3, _vp(c1), cx.sp))
    torch.cuda.synchronize()
    ms_first_query = 1e3 * (time.perf_counter() - t0)
    out["set_ops"] = {
        "cov_ms": ms_cov, "cov": int(cov.value), "cov_GBps": n5 * 8 / (ms_cov * 1e-3) / 1e9,
        "merge_overlaps_ms": ms_merge, "merged_runs": merged_runs, "cov_after_merge_equal": int(cov2.value) == int(cov.value),
        "first_count_after_merge_ms": ms_first_query,
        "union_and_intersect_ms_vs_10M_set": ms_ui, "union": int(u.value), "intersect": int(i.value),
        "note": "wall clock per call through the C ABI on every rank's replica, max over ranks, best of 3 (merge: fresh replica each "
                "time); the merged set's query tables are built on its first count / find (first_count_after_merge_ms)",
    }
    lib.lapper_free(h5)
    lib.lapper_free(h5b)
    return out


def group_selfcheck(cx):
    """lapper_replicate + lapper_group_count/find_batch over the GPUs this process can see (rank 0 only), against the
    single-GPU answer of the same batch."""
    import numpy as np
    import torch

    from rust_lapper_b200 import Lapper, LapperGroup, synth

    ndev = torch.cuda.device_count()
    out = None
    if cx.rank == 0:
        if ndev < 2:
            out = {"devices": ndev, "skipped": "the process sees one GPU"}
        else:
            try:
                devices = list(range(ndev))
                s, e = synth.intervals_gene_exon(SEED_IV3, 2_000_000, UNIVERSE)
                qs, qe = synth.queries_fixed_len(SEED_Q, 4_000_003, UNIVERSE, QLEN)
                qs[::1000] = 0xFFFFFFF0  # a few far-out and inverted queries
                lap = Lapper(starts=s, stops=e, device=cx.local_rank)
                want_c = lap.count_batch(qs, qe)
                want_o, want_i = lap.find_batch(qs, qe)
                t0 = time.perf_counter()
                grp = LapperGroup(lap, devices)
                rep_ms = 1e3 * (time.perf_counter() - t0)
                got_c = grp.count_batch(qs, qe)
                t0 = time.perf_counter()
                got_o, got_i = grp.find_batch(qs, qe)
                find_ms = 1e3 * (time.perf_counter() - t0)
                out = {
                    "devices": devices, "intervals": len(s), "queries": len(qs),
                    "replicate_ms": rep_ms, "group_find_ms": find_ms,
                    "count_equal": bool(np.array_equal(want_c, got_c)),
                    "find_equal": bool(np.array_equal(want_o, got_o) and np.array_equal(want_i, got_i)),
                    "hits": int(want_o[-1]),
                    "note": "rank 0 alone drives every GPU of the box through lapper_group_*; replicate_ms includes creating this "
                            "process's CUDA context and memory pool on each peer (about 0.4 s per GPU, once per process)",
                }
                grp.close()
                lap.close()
            except Exception as ex:
                out = {"devices": ndev, "error": repr(ex)}
    # the other ranks wait on the HOST (rendezvous store), not in an NCCL kernel spinning on the GPUs rank 0 is using
    if cx.dist is not None:
        store = cx.dist.distributed_c10d._get_default_store()
        if cx.rank == 0:
            store.set("lapper_selfcheck_done", "1")
        else:
            store.wait(["lapper_selfcheck_done"])
    cx.barrier()
    return out


def run_extras(cx, n, nq):
    """Other shapes of the same hot path, reported beside the headline (never instead of it)."""
    import torch

    from rust_lapper_b200 import _ffi, synth

    args, lib, dev, stream, sp, peak = cx.args, cx.lib, cx.dev, cx.stream, cx.sp, cx.peak
    out = {}

    def timed(fn, steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    iv_s, iv_e = synth.intervals_uniform(SEED_IV, n, UNIVERSE, 50, 5000, device=dev)
    h = C.c_void_p(0)
    _ffi.check(lib.lapper_build_u32_dev(_vp(iv_s), _vp(iv_e), n, dev.index, 0, sp, C.byref(h)))
    del iv_s, iv_e

    # variant B: the same count kernel on start-sorted queries
    qs, qe = synth.queries_sorted_fixed_len(45, nq, UNIVERSE, QLEN, device=dev)
    counts = torch.empty(nq, dtype=torch.int32, device=dev)
    ms = timed(lambda: _ffi.check(lib.lapper_count_batch_u32_dev(h, _vp(qs), _vp(qe), nq, _vp(counts), sp)), 10)
    gbs = (nq * 12 + n * 8) / (ms * 1e-3) / 1e9
    out["count_sorted_queries"] = {"queries_per_s": nq / (ms * 1e-3), "ms": ms, "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
    del qs, qe

    # the headline batch from PAGEABLE host arrays (what a Vec<u32> / numpy caller passes): the library stages it through
    # its own pinned buffers with a few copy threads (api.cu count_host_staged)
    import numpy as np

    pq_s, pq_e = synth.queries_fixed_len(SEED_Q, nq, UNIVERSE, QLEN)  # numpy, pageable
    p_out = np.zeros(nq, np.uint32)
    npp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    _ffi.check(lib.lapper_count_batch_u32_c32(h, npp(pq_s), npp(pq_e), nq, npp(p_out)))
    t0 = time.perf_counter()
    for _ in range(2):
        _ffi.check(lib.lapper_count_batch_u32_c32(h, npp(pq_s), npp(pq_e), nq, npp(p_out)))
    pg_s = (time.perf_counter() - t0) / 2
    out["e2e_pageable_host_buffers"] = {
        "ms": 1e3 * pg_s, "queries_per_s": nq / pg_s, "counts_sum": int(p_out.sum(dtype=np.uint64)),
        "equals_headline_checksum": int(p_out.sum(dtype=np.uint64)) == cx.headline_checksum,
        "copy_threads": int(os.environ.get("LAPPER_HOST_COPY_THREADS", 4)),
        "note": "lapper_count_batch_u32_c32 with unpinned numpy arrays; the driver's own pageable path (LAPPER_HOST_COPY_THREADS=0) "
                "takes 94 ms for the same call on the bench boxes, pinned buffers 17 ms (the `e2e` block)",
    }
    del pq_s, pq_e, p_out

    # longer reads, random order: the default packed table (margin for 150-bp queries) and the table after
    # lapper_tune_query_length (margin covering the announced length); the handle is tuned back afterwards
    longer = {}
    for ql in (250, 400):
        lq_s, lq_e = synth.queries_fixed_len(SEED_Q + ql, nq, UNIVERSE, ql, device=dev)
        ms_d = timed(lambda: _ffi.check(lib.lapper_count_batch_u32_dev(h, _vp(lq_s), _vp(lq_e), nq, _vp(counts), sp)), 5)
        sum_d = int(counts.to(torch.int64).sum().item())
        _ffi.check(lib.lapper_tune_query_length(h, ql))
        ms_t = timed(lambda: _ffi.check(lib.lapper_count_batch_u32_dev(h, _vp(lq_s), _vp(lq_e), nq, _vp(counts), sp)), 5)
        sum_t = int(counts.to(torch.int64).sum().item())
        _ffi.check(lib.lapper_tune_query_length(h, 0))
        longer[f"{ql}bp"] = {"default_table_ms": ms_d, "tuned_table_ms": ms_t, "tuned_queries_per_s": nq / (ms_t * 1e-3),
                             "same_counts": sum_d == sum_t}
        del lq_s, lq_e
    out["count_random_longer_queries"] = dict(longer, note="lapper_tune_query_length(h, len): the packed count table rebuilt with a "
                                              "margin that covers the announced query length; 100 M random-order queries per launch")

    # variant C: random-order queries INCLUDING an on-device sort by start, the count
    # on the sorted order, and the un-permute of the results.  The sort here is torch.sort (a library call) used
    # only to quantify this alternative -- the engine itself never pre-sorts, because this is slower than
    # counting the random order directly.
    rq_s, rq_e = synth.queries_fixed_len(SEED_Q, nq, UNIVERSE, QLEN, device=dev)
    out_c = torch.empty(nq, dtype=torch.int32, device=dev)

    def variant_c():
        key = rq_s.to(torch.int64) & 0xFFFFFFFF
        _, order = torch.sort(key)
        ss, se = rq_s[order].contiguous(), rq_e[order].contiguous()
        _ffi.check(lib.lapper_count_batch_u32_dev(h, _vp(ss), _vp(se), nq, _vp(counts), sp))
        out_c[order] = counts

    ms_c = timed(variant_c, 3)
    out["count_random_incl_device_sort"] = {
        "queries_per_s": nq / (ms_c * 1e-3), "ms": ms_c,
        "note": "torch.sort + gather + count on sorted order + scatter back; direct random-order count is the headline `value`",
    }
    del rq_s, rq_e, out_c, counts
    lib.lapper_free(h)
    torch.cuda.empty_cache()

    # cfg 4: pathological set (1 % of the intervals span > 50 % of the universe; max_len ~ 0.9 U).  count is
    # unaffected (BITS); find returns ~1 % of all intervals per query, so the batch is small and output-bound.
    s4, e4 = synth.intervals_pathological(46, n, UNIVERSE, device=dev)
    h4 = C.c_void_p(0)
    _ffi.check(lib.lapper_build_u32_dev(_vp(s4), _vp(e4), n, dev.index, 0, sp, C.byref(h4)))
    q4s, q4e = synth.queries_fixed_len(47, nq, UNIVERSE, QLEN, device=dev)
    c4 = torch.empty(nq, dtype=torch.int32, device=dev)
    ms = timed(lambda: _ffi.check(lib.lapper_count_batch_u32_dev(h4, _vp(q4s), _vp(q4e), nq, _vp(c4), sp)), 5)
    nq4 = 10_000
    off4 = torch.empty(nq4 + 1, dtype=torch.int64, device=dev)
    t4 = C.c_uint64(0)
    _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h4, _vp(q4s), _vp(q4e), nq4, _vp(off4), C.byref(t4), sp))
    idx4 = torch.empty(max(int(t4.value), 1), dtype=torch.int32, device=dev)

    def find4():
        t = C.c_uint64(0)
        _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h4, _vp(q4s), _vp(q4e), nq4, _vp(off4), C.byref(t), sp))
        _ffi.check(lib.lapper_find_batch_fill_u32_dev(h4, _vp(q4s), _vp(q4e), nq4, _vp(off4), _vp(idx4), sp))

    ms_f = timed(find4, 3)
    out["cfg4_pathological_long_intervals"] = {
        "max_len": int(lib.lapper_max_len(h4)),
        "count_random_100M_ms": ms, "count_queries_per_s": nq / (ms * 1e-3),
        "find_queries": nq4, "find_hits": int(t4.value), "find_ms": ms_f,
        "find_hits_per_s": int(t4.value) / (ms_f * 1e-3), "find_output_GBps": int(t4.value) * 4 / (ms_f * 1e-3) / 1e9,
        "note": "the reference scans ~n/2 intervals per find() here (src/lib.rs:36-47); length classes keep the scan to the long class",
    }
    lib.lapper_free(h4)
    del s4, e4, q4s, q4e, c4, off4, idx4

    # cfg 1a: README interval_bakeoff shape (README.md:51-61): 50,000 intervals, universe 100,000, len 500..80,000,
    # plus one universe-spanning interval; A-vs-A: find and count over the set's own intervals
    (sa, ea), _ = synth.bakeoff_sets(1, 2, 50_000, 100_000)
    ta = torch.from_numpy(sa.view("int32")).to(dev)
    tb = torch.from_numpy(ea.view("int32")).to(dev)
    h1 = C.c_void_p(0)
    _ffi.check(lib.lapper_build_u32_dev(_vp(ta), _vp(tb), ta.numel(), dev.index, 0, sp, C.byref(h1)))
    n1 = ta.numel()
    off1 = torch.empty(n1 + 1, dtype=torch.int64, device=dev)
    t1 = C.c_uint64(0)
    _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h1, _vp(ta), _vp(tb), n1, _vp(off1), C.byref(t1), sp))
    idx1 = torch.empty(max(int(t1.value), 1), dtype=torch.int32, device=dev)
    c1 = torch.empty(n1, dtype=torch.int32, device=dev)

    def find1():
        t = C.c_uint64(0)
        _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h1, _vp(ta), _vp(tb), n1, _vp(off1), C.byref(t), sp))
        _ffi.check(lib.lapper_find_batch_fill_u32_dev(h1, _vp(ta), _vp(tb), n1, _vp(off1), _vp(idx1), sp))

    ms_f1 = timed(find1, 3)
    ms_c1 = timed(lambda: _ffi.check(lib.lapper_count_batch_u32_dev(h1, _vp(ta), _vp(tb), n1, _vp(c1), sp)), 10)
    out["cfg1a_readme_bakeoff_A_vs_A"] = {
        "intervals": n1, "hits": int(t1.value), "find_ms": ms_f1, "find_hits_per_s": int(t1.value) / (ms_f1 * 1e-3),
        "count_ms": ms_c1, "count_equals_find_total": int(c1.to(torch.int64).sum().item()) == int(t1.value),
        "published_reference": "README.md:73-76: find 4.78125 s (1,469,068,763 hits, unseeded data), count 15.625 ms (timer floor), unstated CPU",
    }
    lib.lapper_free(h1)
    del idx1, off1, ta, tb, c1
    torch.cuda.empty_cache()

    # Lapper<u64, T> (the README's usize) on the cfg-2 shape: the coordinates fit 32 bits, so the u32 engine answers
    # (narrowed index, queries clamped on the way in); the 64-bit kernels (bisection) timed beside it on a tenth
    nq64 = min(nq, 100_000_000)
    iv_s, iv_e = synth.intervals_uniform(SEED_IV, n, UNIVERSE, 50, 5000, device=dev)
    s64 = (iv_s.to(torch.int64) & 0xFFFFFFFF).contiguous()
    e64 = (iv_e.to(torch.int64) & 0xFFFFFFFF).contiguous()
    del iv_s, iv_e
    h64 = C.c_void_p(0)
    _ffi.check(lib.lapper64_build_dev(_vp(s64), _vp(e64), n, dev.index, sp, C.byref(h64)))
    # the same set translated by 2^40 (timestamps rather than a genome): beyond 32 bits, but the span fits, so the u32
    # engine answers on coordinates relative to the smallest one (csrc/engine64.cu, "rebased form")
    shift64 = 1 << 40
    s64 += shift64
    e64 += shift64
    h64r = C.c_void_p(0)
    _ffi.check(lib.lapper64_build_dev(_vp(s64), _vp(e64), n, dev.index, sp, C.byref(h64r)))
    del s64, e64
    rq_s, rq_e = synth.queries_fixed_len(SEED_Q, nq64, UNIVERSE, QLEN, device=dev)
    q64s = (rq_s.to(torch.int64) & 0xFFFFFFFF).contiguous()
    q64e = (rq_e.to(torch.int64) & 0xFFFFFFFF).contiguous()
    c64 = torch.empty(nq64, dtype=torch.int64, device=dev)
    ms_n = timed(lambda: _ffi.check(lib.lapper64_count_batch_dev(h64, _vp(q64s), _vp(q64e), nq64, _vp(c64), sp)), 5)
    narrow = bool(lib.lapper64_is_narrow(h64))
    sum_n = int(c64.sum().item())
    old_mode = lib.lapper64_debug_set_narrow_mode(0)
    nqw = nq64 // 10
    ms_w = timed(lambda: _ffi.check(lib.lapper64_count_batch_dev(h64, _vp(q64s), _vp(q64e), nqw, _vp(c64), sp)), 3)
    lib.lapper64_debug_set_narrow_mode(old_mode)
    q64s += shift64
    q64e += shift64
    ms_r = timed(lambda: _ffi.check(lib.lapper64_count_batch_dev(h64r, _vp(q64s), _vp(q64e), nq64, _vp(c64), sp)), 5)
    rebased = bool(lib.lapper64_is_narrow(h64r))
    sum_r = int(c64.sum().item())
    lib.lapper64_free(h64r)
    out["u64_coordinates_cfg2"] = {
        "count_random_ms": ms_n, "queries": nq64, "queries_per_s": nq64 / (ms_n * 1e-3), "answered_by_u32_engine": narrow,
        "counts_sum": sum_n, "equals_u32_headline_checksum": (sum_n == cx.headline_checksum) if nq64 == nq else None,
        "wide_kernels_ms": ms_w, "wide_kernels_queries": nqw, "wide_kernels_queries_per_s": nqw / (ms_w * 1e-3),
        "translated_by_2^40": {"count_random_ms": ms_r, "queries_per_s": nq64 / (ms_r * 1e-3), "answered_by_u32_engine": rebased,
                               "same_counts_sum": sum_r == sum_n},
        "note": "lapper64_count_batch_dev, u64 queries and u64 counts on the device; 28 B/query instead of 12: the narrowing pass "
                "reads 16 B and writes 8 B per query in front of the u32 kernel",
    }
    lib.lapper64_free(h64)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--intervals", type=int, default=N_INTERVALS)
    ap.add_argument("--queries", type=int, default=NQ_PER_GPU, help="queries per GPU per step")
    ap.add_argument("--find-queries", type=int, default=100_000_000)
    ap.add_argument("--big-intervals", type=int, default=CFG5_INTERVALS, help="cfg-5 leg: intervals (0 = skip)")
    ap.add_argument("--big-queries", type=int, default=CFG5_QUERIES, help="cfg-5 leg: queries in all (sharded over the ranks)")
    ap.add_argument("--big-find-queries", type=int, default=CFG5_FIND_QUERIES, help="cfg-5 leg: find queries in all (0 = skip)")
    ap.add_argument("--no-find", action="store_true", help="skip the cfg-3 find leg")
    ap.add_argument("--no-e2e-find", action="store_true", help="skip the host-pointer find leg (pins ~5.6 GB)")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the lapper_group_* self-check over peer GPUs")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra shapes (N = 1 only)")
    ap.add_argument("--extras", action="store_true", help=argparse.SUPPRESS)  # kept for old command lines; extras are on by default
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
