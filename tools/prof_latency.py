"""Per-call latency of the single-query facade path (what a loop over Lapper::count / Lapper::find costs): batches of
one query through lapper_count_batch_u32 / lapper_find_batch_u32 + lapper_result_copy, host pointers, ctypes included."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rust_lapper_b200 import Lapper, _ffi  # noqa: E402

n = int(os.environ.get("N_IV", 100_000))
reps = int(os.environ.get("REPS", 3000))
rng = np.random.default_rng(1)
s = rng.integers(0, 3_000_000_000, n, dtype=np.uint32)
lap = Lapper(starts=s, stops=s + rng.integers(50, 5000, n).astype(np.uint32))
lib = _ffi.lib()
qs = rng.integers(0, 3_000_000_000, reps, dtype=np.uint32)
for name, fn in (("count", lambda a: lap.count(int(a), int(a) + 150)), ("find", lambda a: lap.find_idx(int(a), int(a) + 150))):
    for a in qs[:200]:
        fn(a)
    t0 = time.perf_counter()
    for a in qs:
        fn(a)
    dt = time.perf_counter() - t0
    print(f"{name}: {dt / reps * 1e6:.1f} us per call ({reps} calls, n = {n})")
for m in (100, 10_000):
    a = qs[:m] if m <= reps else rng.integers(0, 3_000_000_000, m, dtype=np.uint32)
    b = a + np.uint32(150)
    lap.count_batch(a, b)
    t0 = time.perf_counter()
    for _ in range(200):
        lap.count_batch(a, b)
    print(f"count_batch of {m}: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per call")
