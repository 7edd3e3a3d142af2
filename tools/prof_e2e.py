"""End-to-end count from pinned host memory (lapper_count_batch_u32_c32): H2D + kernels + D2H inside the call."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rust_lapper_b200 import _ffi, synth  # noqa: E402

n = int(os.environ.get("N_IV", 10_000_000))
nq = int(os.environ.get("NQ", 100_000_000))
reps = int(os.environ.get("REPS", 5))
dev = torch.device("cuda", 0)
lib = _ffi.lib()
s, e = synth.intervals_uniform(42, n, synth.GRCH38, 50, 5000, device=dev)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
h = C.c_void_p(0)
_ffi.check(lib.lapper_build_u32_dev(C.c_void_p(s.data_ptr()), C.c_void_p(e.data_ptr()), n, 0, 0, sp, C.byref(h)))
qs, qe = synth.queries_fixed_len(43, nq, synth.GRCH38, 150, device=dev)
hs = torch.empty(nq, dtype=torch.int32, pin_memory=True)
he = torch.empty(nq, dtype=torch.int32, pin_memory=True)
ho = torch.empty(nq, dtype=torch.int32, pin_memory=True)
hs.copy_(qs)
he.copy_(qe)
torch.cuda.synchronize()
vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
for _ in range(2):
    _ffi.check(lib.lapper_count_batch_u32_c32(h, vp(hs), vp(he), nq, vp(ho)))
t0 = time.perf_counter()
for _ in range(reps):
    _ffi.check(lib.lapper_count_batch_u32_c32(h, vp(hs), vp(he), nq, vp(ho)))
ms = (time.perf_counter() - t0) / reps * 1e3
print(f"e2e count: {ms:.3f} ms per call, {nq / ms / 1e6:.2f} Gq/s, {(nq * 12) / ms / 1e6:.1f} GB/s over PCIe (both directions), checksum {int(ho.to(torch.int64).sum())}")
lib.lapper_free(h)
