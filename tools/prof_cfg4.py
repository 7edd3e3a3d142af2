"""cfg 4 (pathological long intervals): find on 10 k random queries, device-resident (heavy-query path)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rust_lapper_b200 import _ffi, synth  # noqa: E402

n = int(os.environ.get("N_IV", 10_000_000))
nq = int(os.environ.get("NQ", 10_000))
reps = int(os.environ.get("REPS", 5))
dev = torch.device("cuda", 0)
lib = _ffi.lib()
s, e = synth.intervals_pathological(46, n, synth.GRCH38, device=dev)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
h = C.c_void_p(0)
_ffi.check(lib.lapper_build_u32_dev(C.c_void_p(s.data_ptr()), C.c_void_p(e.data_ptr()), n, 0, 0, sp, C.byref(h)))
qs, qe = synth.queries_fixed_len(47, nq, synth.GRCH38, 150, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
off = torch.empty(nq + 1, dtype=torch.int64, device=dev)
total = C.c_uint64(0)
_ffi.check(lib.lapper_find_batch_offsets_u32_dev(h, vp(qs), vp(qe), nq, vp(off), C.byref(total), sp))
idx = torch.empty(max(int(total.value), 1), dtype=torch.int32, device=dev)


def step():
    t = C.c_uint64(0)
    _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h, vp(qs), vp(qe), nq, vp(off), C.byref(t), sp))
    _ffi.check(lib.lapper_find_batch_fill_u32_dev(h, vp(qs), vp(qe), nq, vp(off), vp(idx), sp))


for _ in range(2):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"cfg4 find: {ms:.3f} ms per pass, {int(total.value)} hits, {int(total.value) / ms / 1e6:.1f} Ghits/s, "
      f"checksum {int(idx[: int(total.value)].to(torch.int64).sum())}")
lib.lapper_free(h)
