"""Randomised parity fuzz: the engine (u32 and u64, every table / kernel path it can be pushed onto) against the CPU
oracle on seeded random cases, for a wall-clock budget (tests/test_gpu_fuzz.py runs it briefly in the suite; run it
longer by hand on a GPU box):

    python tests/fuzz_parity.py --seconds 240 --seed 1

Every case draws: the interval set (size, universe, length law incl. zero-length / duplicates / a few huge / optionally
inverted intervals, coordinates up to the top of the type), the build flags (auto, sector table at several shifts, packed
table forced, no tables), the query batch (size around the kernels' tile / run / threshold boundaries; random, sorted,
sorted-by-start-only, block-shuffled; fixed or random length; empty, inverted and MAX-coordinate queries mixed in), and
checks count_batch (u64 and c32), find_batch (CSR byte-equal), find_batch_into (capacity protocol) and, now and then,
merge_overlaps / cov, union_and_intersect against a second random set, depth, the seek cursor protocol, insert.  u64
cases also draw which engine answers a set that fits 32 bits (the 64-bit kernels / the default policy / the u32 engine
from the first batch on).  Prints one line per failure with everything needed to replay it, and a summary."""
import argparse
import sys
import time
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lapper_oracle  # noqa: E402
from rust_lapper_b200 import Lapper, Lapper64, LapperError, LapperGroup, _ffi  # noqa: E402


def draw_intervals(rng, bits):
    top = (1 << bits) - 1
    n = int(rng.choice([0, 1, 2, 7, 33, 300, 3000, 20_000, 120_000], p=[.02, .03, .05, .1, .15, .25, .2, .15, .05]))
    kind = rng.choice(["dense", "sparse", "top", "clustered"])
    if kind == "dense":
        U = int(rng.integers(50, 5000))
    elif kind == "sparse":
        U = int(rng.integers(1 << 20, min(top, 1 << 40), dtype=np.uint64))
    elif kind == "top":
        U = top
    else:
        U = int(rng.integers(10_000, 5_000_000))
    dt = np.uint64
    if kind == "top":
        s = (np.uint64(top) - rng.integers(0, 1 << 22, n).astype(dt)) if rng.random() < 0.5 else rng.integers(0, np.uint64(top), n, dtype=dt, endpoint=True)
    elif kind == "clustered":
        centers = rng.integers(0, U, max(1, n // 200 + 1), dtype=dt)
        s = centers[rng.integers(0, centers.size, n)] + rng.integers(0, 50, n).astype(dt)
    else:
        s = rng.integers(0, U, n, dtype=dt)
    law = rng.choice(["short", "mixed", "zero", "long"])
    if law == "short":
        ln = rng.integers(0, 40, n).astype(dt)
    elif law == "mixed":
        ln = rng.integers(0, max(2, U // 50 + 2), n, dtype=dt)
    elif law == "zero":
        ln = (rng.integers(0, 3, n) * rng.integers(0, 30, n)).astype(dt)
    else:
        ln = rng.integers(0, 200, n).astype(dt)
        k = rng.random(n) < 0.01
        ln[k] = rng.integers(U // 4 + 1, U // 2 + 2, int(k.sum()), dtype=dt)
    e = np.minimum(s.astype(object) + ln.astype(object), top).astype(dt) if n else s.copy()
    inverted = False
    if n and rng.random() < 0.12:
        k = rng.random(n) < 0.05
        s[k], e[k] = e[k], s[k]
        inverted = bool((e < s).any())
    if n and rng.random() < 0.3:  # duplicates
        k = rng.integers(0, n, n // 3)
        s[: k.size], e[: k.size] = s[k], e[k]
        inverted = bool((e < s).any())
    return s, e, U, inverted


def draw_queries(rng, bits, U, s):
    top = (1 << bits) - 1
    nq = int(rng.choice([0, 1, 31, 257, 1023, 1025, 4099, 33_000, 66_001, 140_000], p=[.02, .05, .08, .1, .1, .1, .15, .2, .15, .05]))
    dt = np.uint64
    span = min(top, U + U // 4 + 10)
    src = rng.choice(["uniform", "near_intervals"])
    if src == "near_intervals" and s.size:
        a = s[rng.integers(0, s.size, nq)].astype(object) + rng.integers(-60, 60, nq).astype(object)
        a = np.clip(a, 0, top).astype(dt)
    else:
        a = rng.integers(0, np.uint64(span), nq, dtype=dt, endpoint=True)
    ql = rng.choice(["fixed", "random", "zero"])
    if ql == "fixed":
        ln = np.full(nq, int(rng.choice([1, 37, 150, 1000])), dtype=dt)
    elif ql == "random":
        ln = rng.integers(0, max(2, min(U // 20 + 2, 1 << 30)), nq).astype(dt)
    else:
        ln = (rng.integers(0, 2, nq) * rng.integers(0, 200, nq)).astype(dt)
    b = np.minimum(a.astype(object) + ln.astype(object), top).astype(dt) if nq else a.copy()
    order = rng.choice(["random", "sorted", "sorted_start", "blocks"])
    if order in ("sorted", "sorted_start", "blocks") and nq:
        o = np.argsort(a, kind="stable")
        a, b = a[o], b[o]
        if order == "sorted":
            b = np.maximum.accumulate(b)
        if order == "blocks":  # sorted runs in shuffled order
            k = 256 * int(rng.integers(1, 6))
            blocks = [np.arange(i, min(i + k, nq)) for i in range(0, nq, k)]
            rng.shuffle(blocks)
            p = np.concatenate(blocks)
            a, b = a[p], b[p]
    if nq and rng.random() < 0.5:  # special queries: inverted, empty, at the ends
        k = rng.integers(0, nq, max(1, nq // 50))
        a[k], b[k] = b[k], a[k]
        m = min(nq, 4)
        a[:m] = np.array([0, top, top, 0], dtype=dt)[:m]
        b[:m] = np.array([0, top, 0, top], dtype=dt)[:m]
        if bits == 64 and nq > 40:  # coordinates around 2^32: where a narrowed u64 set clamps its queries
            t32 = (1 << 32) - 1
            pts = np.array([t32 - 3, t32 - 2, t32 - 1, t32, t32 + 1, 1 << 33, top - 1, top], dtype=dt)
            k = rng.integers(4, nq, 32)
            a[k] = pts[rng.integers(0, pts.size, 32)]
            b[k] = pts[rng.integers(0, pts.size, 32)]
    return a, b, order


def check_case(rng, bits, case_id, verbose=False):
    s, e, U, inverted = draw_intervals(rng, bits)
    qs, qe, order = draw_queries(rng, bits, U, s)
    desc = f"bits={bits} n={s.size} U={U} inverted={inverted} nq={qs.size} order={order}"
    if bits == 32:
        flags = int(rng.choice([0, 0, 1, 2, 2 | ((0 + 1) << 8), (3 + 1) << 8, 8 | 2, 8 | 2 | (int(rng.integers(2, 1800)) << 16)]))
        desc += f" flags={flags}"
        try:
            gpu = Lapper(starts=s.astype(np.uint32), stops=e.astype(np.uint32), flags=flags)
        except LapperError as ex:  # a forced table shape that the set cannot take is an error, not a wrong answer
            if ex.code == _ffi.ERR_INVALID_ARG:
                return None
            raise
        cpu = lapper_oracle.OracleLapper(starts=s.astype(np.uint32), stops=e.astype(np.uint32))
        qs, qe = qs.astype(np.uint32), qe.astype(np.uint32)
        if rng.random() < 0.2:  # an announced query length reshapes the packed table, never the answers
            ql = int(rng.choice([0, 1, 37, 150, 300, 640, 641, 5000]))
            gpu.tune_query_length(ql)
            desc += f" tuned={ql}"
    else:
        # which engine answers a u64 set whose coordinates fit 32 bits: the 64-bit kernels, the default policy (u32
        # engine from the first large batch on), or the u32 engine from the start
        mode = int(rng.choice([0, 1, 2, 2]))
        _ffi.lib().lapper64_debug_set_narrow_mode(mode)
        desc += f" narrow_mode={mode}"
        top = (1 << 64) - 1
        hi = int(max(s.max(), e.max())) if s.size else top
        if top - hi > (1 << 33) and rng.random() < 0.35:
            # the whole case translated beyond 32 bits (now and then right up to u64::MAX): a set whose span still fits
            # 32 bits is answered by the u32 engine on rebased coordinates (csrc/engine64.cu, "rebased form")
            off = top - hi if rng.random() < 0.2 else int(rng.integers(1 << 32, min(top - hi, 1 << 63), dtype=np.uint64))
            s, e = s + np.uint64(off), e + np.uint64(off)
            qs = np.minimum(qs.astype(object) + off, top).astype(np.uint64)
            qe = np.minimum(qe.astype(object) + off, top).astype(np.uint64)
            lo, hi = int(min(s.min(), e.min())), int(max(s.max(), e.max()))
            if qs.size > 40:
                pts = np.array([0, 1, lo - 1, lo, lo + 1, hi - 1, hi, min(hi + 1, top), top - 1, top], dtype=np.uint64)
                k = rng.integers(4, qs.size, 32)
                qs[k] = pts[rng.integers(0, pts.size, 32)]
                qe[k] = pts[rng.integers(0, pts.size, 32)]
            desc += f" shifted_by={off}"
        gpu = Lapper64(starts=s, stops=e)
        cpu = lapper_oracle.OracleLapper(starts=s, stops=e, coord_bits=64)
    want_c = cpu.count_batch(qs, qe, 4)
    got_c = gpu.count_batch(qs, qe)
    if not np.array_equal(got_c, want_c):
        bad = np.nonzero(got_c != want_c)[0]
        if verbose:
            for j in bad[:8].tolist():
                print(f"  query {j}: ({int(qs[j])}, {int(qe[j])}) want {int(want_c[j])} got {int(got_c[j])}")
            if bits == 32:
                print(f"  bucket_shift {gpu.bucket_shift} packed {gpu.packed_table} max_len {gpu.max_len} n {len(gpu)}; "
                      f"starts [{int(s.min())}, {int(s.max())}] stops [{int(e.min())}, {int(e.max())}]")
        return f"count_batch differs at {bad[:5].tolist()} ({bad.size} queries): {desc}"
    if bits == 32 and not np.array_equal(gpu.count_batch_c32(qs, qe), want_c.astype(np.uint32)):
        return f"count_batch_c32 differs: {desc}"
    wo, wi = cpu.find_batch(qs, qe, 4)
    if int(wo[-1]) < 60_000_000:
        go, gi = gpu.find_batch(qs, qe)
        if not (np.array_equal(go, wo) and np.array_equal(gi, wi)):
            return f"find_batch differs: {desc}"
        if rng.random() < 0.5:  # a second batch on the same handle: find_batch sizes its buffer from the first one
            go, gi = gpu.find_batch(qs[::-1].copy(), qe[::-1].copy())
            wo2, wi2 = cpu.find_batch(qs[::-1].copy(), qe[::-1].copy(), 4)
            if not (np.array_equal(go, wo2) and np.array_equal(gi, wi2)):
                return f"second find_batch differs: {desc}"
        if bits == 32 and rng.random() < 0.3:
            off = np.empty(qs.size + 1, np.uint64)
            idx = np.empty(int(wo[-1]), np.uint32)
            t = gpu.find_batch_into(qs, qe, off, idx)
            if t != int(wo[-1]) or not (np.array_equal(off, wo) and np.array_equal(idx, wi)):
                return f"find_batch_into differs: {desc}"
    if bits == 32 and s.size and rng.random() < 0.05 and int(wo[-1]) < 60_000_000:
        # replica group on this GPU (three replicas of the handle; the batch is cut into three shards)
        grp = LapperGroup(gpu, [0, 0, 0])
        try:
            if not np.array_equal(grp.count_batch(qs, qe), want_c):
                return f"group count_batch differs: {desc}"
            go, gi = grp.find_batch(qs, qe)
            if not (np.array_equal(go, wo) and np.array_equal(gi, wi)):
                return f"group find_batch differs: {desc}"
        finally:
            grp.close()
    # the rest of the API now and then, on sets small enough for the per-call oracle: union_and_intersect against a
    # second random set, depth, the seek cursor protocol over a sorted batch, insert (positions, perm, max_len)
    if not inverted and 0 < s.size <= 20_000 and rng.random() < 0.25:
        s2, e2, _, inv2 = draw_intervals(rng, bits)
        if not inv2 and 0 < s2.size <= 5000:
            if bits == 32:
                g2 = Lapper(starts=s2.astype(np.uint32), stops=e2.astype(np.uint32))
                c2 = lapper_oracle.OracleLapper(starts=s2.astype(np.uint32), stops=e2.astype(np.uint32))
            else:
                g2 = Lapper64(starts=s2, stops=e2)
                c2 = lapper_oracle.OracleLapper(starts=s2, stops=e2, coord_bits=64)
            if gpu.union_and_intersect(g2) != cpu.union_and_intersect(c2) or g2.union_and_intersect(gpu) != c2.union_and_intersect(cpu):
                return f"union_and_intersect differs (second set n={s2.size}): {desc}"
        if s.size <= 5000:
            # (the reference's depth iterator walks the merged intervals coordinate by coordinate, src/lib.rs:744-784, and so
            # does the oracle: only sets that cover little)
            if cpu.cov() <= 100_000 and gpu.depth() != cpu.depth():
                return f"depth differs: {desc}"
            o = np.argsort(qs[:300], kind="stable")
            cg, cc = [0], [0]
            for j in o.tolist():
                if gpu.seek_idx(int(qs[j]), int(qe[j]), cg) != cpu.seek_idx(int(qs[j]), int(qe[j]), cc) or cg != cc:
                    return f"seek differs at query {j}: {desc}"
    if 0 < s.size <= 5000 and rng.random() < 0.2:
        top = (1 << bits) - 1
        for i in range(4):
            k = int(rng.integers(0, s.size))
            a = int(s[k]) if rng.random() < 0.5 else int(rng.integers(0, min(top, 2 * U + 10), dtype=np.uint64))
            b = min(top, a + int(rng.integers(0, 500)))
            gpu.insert(a, b), cpu.insert(a, b, s.size + i)  # (the oracle's val = the input position the engine hands out)
        gs, ge, gp = gpu.intervals_soa()
        cs, ce, cv = cpu.intervals_soa()
        if not (np.array_equal(gs, cs) and np.array_equal(ge, ce) and gpu.max_len == cpu.max_len and np.array_equal(gpu.stops, cpu.stops)):
            return f"insert differs: {desc}"
        if not np.array_equal(gpu.count_batch(qs, qe), cpu.count_batch(qs, qe, 4)):
            return f"count_batch after insert differs: {desc}"
        wo, wi = cpu.find_batch(qs, qe, 4)
        if int(wo[-1]) < 60_000_000:
            go, gi = gpu.find_batch(qs, qe)
            if not (np.array_equal(go, wo) and np.array_equal(gi, wi)):
                return f"find_batch after insert differs: {desc}"
        inverted = bool((ge < gs).any())
    if s.size and rng.random() < 0.3:
        if not inverted and gpu.cov() != cpu.cov():
            return f"cov differs: {desc}"
        gpu.merge_overlaps(), cpu.merge_overlaps()
        gs, ge, gp = gpu.intervals_soa()
        cs, ce, cv = cpu.intervals_soa()
        if not (np.array_equal(gs, cs) and np.array_equal(ge, ce) and np.array_equal(gp, cv)):
            return f"merge_overlaps differs: {desc}"
        if not np.array_equal(gpu.count_batch(qs, qe), cpu.count_batch(qs, qe, 4)):
            return f"count_batch after merge differs: {desc}"
        wo, wi = cpu.find_batch(qs, qe, 4)
        if int(wo[-1]) < 60_000_000:
            go, gi = gpu.find_batch(qs, qe)
            if not (np.array_equal(go, wo) and np.array_equal(gi, wi)):
                return f"find_batch after merge differs: {desc}"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--case", type=int, default=None, help="replay one case_seed verbosely")
    ap.add_argument("--u64-share", type=float, default=0.3, help="share of the cases drawn for the u64 engine")
    args = ap.parse_args()
    if args.case is not None:
        _ffi.lib().lapper_debug_set_sorted_min_queries(32 * 1024)
        rng = np.random.default_rng(args.case)
        bits = 64 if rng.random() < args.u64_share else 32
        print(check_case(rng, bits, 0, verbose=True))
        return
    lib = _ffi.lib()
    lib.lapper_debug_set_sorted_min_queries(32 * 1024)  # offer the sorted-run kernel to oracle-sized batches
    t0 = time.time()
    cases = fails = skipped = 0
    per_bits = {32: 0, 64: 0}
    case_id = 0
    while time.time() - t0 < args.seconds:
        case_seed = args.seed * 1_000_003 + case_id
        rng = np.random.default_rng(case_seed)
        bits = 64 if rng.random() < args.u64_share else 32
        try:
            msg = check_case(rng, bits, case_id)
        except Exception as ex:  # noqa: BLE001
            msg = f"exception {ex!r}"
        if msg is None:
            pass
        if msg:
            fails += 1
            print(f"FAIL case_seed={case_seed}: {msg}", flush=True)
        cases += 1
        per_bits[bits] += 1
        case_id += 1
    stats = (__import__("ctypes").c_uint64 * 2)()
    lib.lapper_debug_sorted_run_stats(stats, 0)
    print(f"fuzz: {cases} cases ({per_bits[32]} u32, {per_bits[64]} u64) in {time.time() - t0:.0f} s, {fails} failures; "
          f"sorted-run kernel: {stats[0]} runs ranked, {stats[1]} per query; debug violations: {lib.lapper_debug_violations()}")
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
