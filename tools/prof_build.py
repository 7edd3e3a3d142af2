"""Profiling driver: index builds (Lapper::new on the device) at a few sizes."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rust_lapper_b200 import _ffi, synth  # noqa: E402

lib = _ffi.lib()
dev = torch.device("cuda", 0)
sizes = [int(x) for x in os.environ.get("SIZES", "1000000,10000000").split(",")]
for n in sizes:
    s, e = synth.intervals_uniform(42, n, synth.GRCH38, 50, 5000, device=dev)
    torch.cuda.synchronize()
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for rep in range(int(os.environ.get("REPS", 3))):
        h = C.c_void_p(0)
        t0 = time.perf_counter()
        _ffi.check(lib.lapper_build_u32_dev(C.c_void_p(s.data_ptr()), C.c_void_p(e.data_ptr()), n, 0, 0, sp, C.byref(h)))
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        lib.lapper_free(h)
        df = time.perf_counter() - t1
        print(f"build n={n} rep {rep}: {dt * 1e3:.2f} ms ({n / dt / 1e6:.0f} M intervals/s), free {df * 1e3:.2f} ms")
