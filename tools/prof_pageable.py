"""count_batch from PAGEABLE host arrays (what a Vec<u32> / numpy caller passes) against pinned ones: the library stages
large pageable batches through its own pinned buffers with a few copy threads (api.cu count_host_staged)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rust_lapper_b200 import _ffi, synth  # noqa: E402

n, nq = 10_000_000, int(os.environ.get("NQ", 100_000_000))
lib = _ffi.lib()
s, e = synth.intervals_uniform(42, n, synth.GRCH38, 50, 5000)
h = C.c_void_p(0)
_ffi.check(lib.lapper_build_u32(s.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), n, 0, C.byref(h)))
qs, qe = synth.queries_fixed_len(43, nq, synth.GRCH38, 150)
out = np.zeros(nq, np.uint32)
vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


def run(a, b, o, reps=3):
    _ffi.check(lib.lapper_count_batch_u32_c32(h, vp(a), vp(b), nq, vp(o)))
    t0 = time.perf_counter()
    for _ in range(reps):
        _ffi.check(lib.lapper_count_batch_u32_c32(h, vp(a), vp(b), nq, vp(o)))
    return (time.perf_counter() - t0) / reps


dt = run(qs, qe, out)
print(f"pageable: {dt * 1e3:.1f} ms per {nq} queries ({nq / dt / 1e9:.2f} Gq/s), checksum {int(out.sum(dtype=np.uint64))}")
pq_s = torch.from_numpy(qs).pin_memory().numpy()
pq_e = torch.from_numpy(qe).pin_memory().numpy()
pout_t = torch.zeros(nq, dtype=torch.int32).pin_memory()
pout = pout_t.numpy().view(np.uint32)
dt = run(pq_s, pq_e, pout)
print(f"pinned:   {dt * 1e3:.1f} ms per {nq} queries ({nq / dt / 1e9:.2f} Gq/s), checksum {int(pout.sum(dtype=np.uint64))}")
assert np.array_equal(out, pout)
lib.lapper_free(h)
