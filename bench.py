#!/usr/bin/env python
"""bench.py -- overlap queries/sec of the batched BITS `count` path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one pass of `count_batch` over one batch of synthetic queries.  Workload (cfg 2 of
BASELINE.json): 10 M intervals (len 50..5000) in a 3.1 Gbp universe x 100 M random-order 150-bp
queries per GPU.  `value` = whole-job queries/s with inputs resident in HBM (one kernel launch per
step, timed with CUDA events); `e2e` = the same through the host-pointer C-ABI call
(lapper_count_batch_u32_c32) with pinned host buffers, H2D + kernel + D2H inside the timed region.
For N > 1 (torchrun, one rank per GPU) the index is replicated and every rank owns its own
contiguous slice of the query stream: weak scaling, no data-path collective.
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
METRIC = "overlap queries/sec (count_batch, BITS)"
UNIT = "queries/s"
WORKLOAD = "cfg2: BITS count, 10M intervals (len 50-5000) in 3.1 Gbp universe x 100M random-order 150-bp queries per GPU"


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
            time.sleep(0.004)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

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


def cpu_count_baseline(n_intervals, sample_q, threads, budget_s=20.0):
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
    # calibrate on 1/16 of the sample, then size the run to the time budget
    cal = max(sample_q // 16, 1)
    t0 = time.perf_counter()
    lap.count_batch(qs[:cal], qe[:cal], threads)
    rate = cal / max(time.perf_counter() - t0, 1e-9)
    m = int(min(sample_q, max(cal, rate * budget_s)))
    t0 = time.perf_counter()
    counts = lap.count_batch(qs[:m], qe[:m], threads)
    dt = time.perf_counter() - t0
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
    }, lap, (qs[:m], qe[:m], counts)


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
    # size one step to roughly 5 s of host work so K steps + W warm-ups end within a few minutes
    t0 = time.perf_counter()
    lap.count_batch(qs[: sample // 20], qe[: sample // 20], threads)
    rate = (sample // 20) / max(time.perf_counter() - t0, 1e-9)
    m = int(min(sample, max(sample // 20, rate * 5.0)))
    for _ in range(args.warmup):
        lap.count_batch(qs[:m], qe[:m], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        counts = lap.count_batch(qs[:m], qe[:m], threads)
    dt = time.perf_counter() - t0
    value = m * args.steps / dt
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
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch

    from rust_lapper_b200 import _ffi, synth

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

    n, nq = args.intervals, args.queries
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    # ---- index: replicated (every rank builds it from the same seeded intervals, on its own GPU)
    iv_s, iv_e = synth.intervals_uniform(SEED_IV, n, UNIVERSE, 50, 5000, device=dev)
    torch.cuda.synchronize()
    h = C.c_void_p(0)
    t0 = time.perf_counter()
    _ffi.check(lib.lapper_build_u32_dev(_vp(iv_s), _vp(iv_e), n, local_rank, 0, sp, C.byref(h)))
    build_ms = 1e3 * (time.perf_counter() - t0)
    shift = int(lib.lapper_bucket_shift(h))

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
    with ClockSampler(local_rank) as clocks:
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

    # per-launch distribution (each launch separately timed) for the roofline of the one kernel
    per_launch = []
    for _ in range(min(args.steps, 10)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step()
        b.record(stream)
        torch.cuda.synchronize()
        per_launch.append(a.elapsed_time(b))
    peak, peak_src = hbm_peak()
    # count_bucket_kernel in its two register builds (one of them exits at once, see csrc/query.cu) + the ragged-tail kernel
    launches_per_step = launches_timed // args.steps
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
                 "about 64 % of those gathers hit L2, the rest cost 64 B of DRAM each (see traffic); L2, L1 and DRAM "
                 "throughput are all at 65-71 % of their peaks in the ncu capture; DESIGN.md section 4")
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

    extras = {}
    if not args.no_extras and world == 1:
        extras = run_extras(args, lib, h, dev, stream, sp, n, nq, peak)

    # ---- CPU baseline beside it (rank 0, N = 1 only) + parity of the timed outputs on the sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu, _, (cq_s, cq_e, ccounts) = cpu_count_baseline(n, min(nq, 30_000_000), os.cpu_count() or 1)
            m = len(cq_s)
            got = counts[:m].cpu().numpy().view(np.uint32).astype(np.uint64)
            cpu["gpu_matches_oracle_on_sample"] = bool(np.array_equal(got, ccounts))
        except Exception as ex:  # the baseline is a report, never a reason to lose the GPU line
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex!r}"}

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
            "config": {
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
                "parallelism": f"index replicated x{world}, queries sharded contiguously, no data-path collective",
                "l2_policy": "inputs (1.2 GB of queries + results per step) are far larger than the 126 MB L2; no flush needed",
            },
            "roofline": roofline,
            "e2e": e2e,
            "gpu_launches": world * launches_timed,
            "gpu_launches_note": "kernels launched by the library inside the timed region, counted by lapper_query_kernel_launches(); "
                                 "per step: count_bucket_kernel at 4 CTAs/SM (runs for start-sorted batches, else exits at once), "
                                 "count_bucket_kernel at 3 CTAs/SM (the reverse), count_bucket_tail_kernel for the ragged tail",
            "clocks": clocks.summary(),
            "checksum": checksum,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extras:
            line["extras"] = extras
        print(json.dumps(line), flush=True)
    lib.lapper_free(h)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_extras(args, lib, h, dev, stream, sp, n, nq, peak):
    """Other shapes of the same hot path, reported beside the headline (never instead of it)."""
    import torch

    from rust_lapper_b200 import _ffi, synth

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

    # variant B: the same count kernel on start-sorted queries
    qs, qe = synth.queries_sorted_fixed_len(45, nq, UNIVERSE, QLEN, device=dev)
    counts = torch.empty(nq, dtype=torch.int32, device=dev)
    ms = timed(lambda: _ffi.check(lib.lapper_count_batch_u32_dev(h, _vp(qs), _vp(qe), nq, _vp(counts), sp)), 10)
    gbs = (nq * 12 + n * 8) / (ms * 1e-3) / 1e9
    out["count_sorted_queries"] = {"queries_per_s": nq / (ms * 1e-3), "ms": ms, "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
    del qs, qe

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

    # cfg 3: find with CSR output, gene/exon-like intervals x start-sorted 150-bp queries
    nq3 = min(nq, args.find_queries)
    s3, e3 = synth.intervals_gene_exon(44, n, UNIVERSE, device=dev)
    h3 = C.c_void_p(0)
    _ffi.check(lib.lapper_build_u32_dev(_vp(s3), _vp(e3), n, dev.index, 0, sp, C.byref(h3)))
    q3s, q3e = synth.queries_sorted_fixed_len(45, nq3, UNIVERSE, QLEN, device=dev)
    offsets = torch.empty(nq3 + 1, dtype=torch.int64, device=dev)
    total = C.c_uint64(0)
    _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h3, _vp(q3s), _vp(q3e), nq3, _vp(offsets), C.byref(total), sp))
    hits = int(total.value)
    indices = torch.empty(max(hits, 1), dtype=torch.int32, device=dev)

    def find_step():
        t = C.c_uint64(0)
        _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h3, _vp(q3s), _vp(q3e), nq3, _vp(offsets), C.byref(t), sp))
        _ffi.check(lib.lapper_find_batch_fill_u32_dev(h3, _vp(q3s), _vp(q3e), nq3, _vp(offsets), _vp(indices), sp))

    ms = timed(find_step, 5)
    fbytes = nq3 * 16 + hits * 4 + n * 8
    gbs = fbytes / (ms * 1e-3) / 1e9
    out["find_cfg3_sorted_queries"] = {
        "queries": nq3, "hits": hits, "hits_per_query": hits / nq3, "queries_per_s": nq3 / (ms * 1e-3),
        "hits_per_s": hits / (ms * 1e-3), "ms": ms, "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peak,
        "max_len": int(lib.lapper_max_len(h3)),
    }
    lib.lapper_free(h3)
    del s3, e3, q3s, q3e, offsets, indices

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
    del idx1, off1

    # cfg 5 shape on one GPU: 50 M intervals (cfg-2 distribution): build, cov, merge_overlaps,
    # union_and_intersect and random-order count against the larger (485 MB) table
    n5 = args.big_intervals
    if n5:
        s5, e5 = synth.intervals_uniform(SEED_IV, n5, UNIVERSE, 50, 5000, device=dev)
        torch.cuda.synchronize()

        def wall(fn, reps=3):
            best = None
            for _ in range(reps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = fn()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            return best * 1e3, r

        def build5():
            hh = C.c_void_p(0)
            _ffi.check(lib.lapper_build_u32_dev(_vp(s5), _vp(e5), n5, dev.index, 0, sp, C.byref(hh)))
            return hh

        handles = []
        ms_build, h5 = wall(lambda: handles.append(build5()) or handles[-1])
        for hh in handles[:-1]:
            lib.lapper_free(hh)
        cov = C.c_uint32(0)
        ms_cov, _ = wall(lambda: _ffi.check(lib.lapper_cov(h5, C.byref(cov))))
        q5s, q5e = synth.queries_fixed_len(SEED_Q, nq, UNIVERSE, QLEN, device=dev)
        c5 = torch.empty(nq, dtype=torch.int32, device=dev)
        ms_cnt = timed(lambda: _ffi.check(lib.lapper_count_batch_u32_dev(h5, _vp(q5s), _vp(q5e), nq, _vp(c5), sp)), 5)
        s5b, e5b = synth.intervals_uniform(SEED_IV + 100, n5 // 5, UNIVERSE, 50, 5000, device=dev)
        h5b = C.c_void_p(0)
        _ffi.check(lib.lapper_build_u32_dev(_vp(s5b), _vp(e5b), n5 // 5, dev.index, 0, sp, C.byref(h5b)))
        u, i = C.c_uint32(0), C.c_uint32(0)
        ms_ui, _ = wall(lambda: _ffi.check(lib.lapper_union_and_intersect(h5, h5b, C.byref(u), C.byref(i))))
        ms_merge, _ = wall(lambda: _ffi.check(lib.lapper_merge_overlaps(h5)), reps=1)
        out["cfg5_50M_intervals_one_gpu"] = {
            "intervals": n5, "build_ms": ms_build, "build_intervals_per_s": n5 / (ms_build * 1e-3),
            "cov_ms": ms_cov, "cov": int(cov.value), "cov_GBps": n5 * 8 / (ms_cov * 1e-3) / 1e9,
            "merge_overlaps_ms": ms_merge, "merged_runs": int(lib.lapper_len(h5)),
            "union_and_intersect_ms_vs_10M_set": ms_ui, "union": int(u.value), "intersect": int(i.value),
            "count_random_100M_ms": ms_cnt, "count_random_queries_per_s": nq / (ms_cnt * 1e-3),
            "note": "wall-clock per call through the C ABI (includes allocations and the D2H of the scalar results)",
        }
        lib.lapper_free(h5)
        lib.lapper_free(h5b)
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
    ap.add_argument("--big-intervals", type=int, default=50_000_000, help="cfg-5 extras leg (0 = skip)")
    ap.add_argument("--no-extras", action="store_true", help="skip the sorted-query count and cfg-3 find legs (N = 1 only)")
    ap.add_argument("--extras", action="store_true", help=argparse.SUPPRESS)  # kept for old command lines; extras are on by default
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
