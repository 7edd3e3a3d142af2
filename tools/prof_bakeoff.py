"""README interval_bakeoff shape (README.md:51-61): 50,000 intervals, universe 100,000, + one universe-spanning interval;
A-vs-A find (every query is heavy: ~29 k hits), device-resident."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rust_lapper_b200 import _ffi, synth  # noqa: E402

reps = int(os.environ.get("REPS", 5))
dev = torch.device("cuda", 0)
lib = _ffi.lib()
(sa, ea), _ = synth.bakeoff_sets(1, 2, 50_000, 100_000)
ta = torch.from_numpy(sa.view("int32")).to(dev)
tb = torch.from_numpy(ea.view("int32")).to(dev)
n = ta.numel()
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
h = C.c_void_p(0)
vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
_ffi.check(lib.lapper_build_u32_dev(vp(ta), vp(tb), n, 0, 0, sp, C.byref(h)))
off = torch.empty(n + 1, dtype=torch.int64, device=dev)
total = C.c_uint64(0)
_ffi.check(lib.lapper_find_batch_offsets_u32_dev(h, vp(ta), vp(tb), n, vp(off), C.byref(total), sp))
idx = torch.empty(max(int(total.value), 1), dtype=torch.int32, device=dev)


def step():
    t = C.c_uint64(0)
    _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h, vp(ta), vp(tb), n, vp(off), C.byref(t), sp))
    _ffi.check(lib.lapper_find_batch_fill_u32_dev(h, vp(ta), vp(tb), n, vp(off), vp(idx), sp))


for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"bakeoff A-vs-A find: {ms:.3f} ms per pass, {int(total.value)} hits, {int(total.value) / ms / 1e6:.1f} Ghits/s, "
      f"{int(total.value) * 4 / ms / 1e6:.0f} GB/s of output, checksum {int(idx[: int(total.value)].to(torch.int64).sum())}")
lib.lapper_free(h)
