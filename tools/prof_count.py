"""Profiling driver: one cfg-2 index, a few count_batch launches on random and on start-sorted
queries (device-resident).  Run plain first, then under ncu (see profiles/README.md)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rust_lapper_b200 import _ffi, synth  # noqa: E402

n = int(os.environ.get("N_IV", 10_000_000))
nq = int(os.environ.get("NQ", 100_000_000))
reps = int(os.environ.get("REPS", 3))
dev = torch.device("cuda", 0)
lib = _ffi.lib()
if os.environ.get("INTERVALS", "uniform") == "gene_exon":  # the cfg-3 interval set
    s, e = synth.intervals_gene_exon(44, n, synth.GRCH38, device=dev)
else:
    s, e = synth.intervals_uniform(42, n, synth.GRCH38, 50, 5000, device=dev)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
h = C.c_void_p(0)
flags = int(os.environ.get("FLAGS", 0))  # e.g. 4 = LAPPER_BUILD_NO_DENSE
_ffi.check(lib.lapper_build_u32_dev(C.c_void_p(s.data_ptr()), C.c_void_p(e.data_ptr()), n, 0, flags, sp, C.byref(h)))
w, ns, no = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
_ffi.check(lib.lapper_packed_table_info(h, C.byref(w), C.byref(ns), C.byref(no)))
print(f"flags {flags}: sector-table shift {lib.lapper_bucket_shift(h)}, packed table width {w.value}, "
      f"{ns.value} sectors ({ns.value * 32 / 1e6:.1f} MB), {no.value} overflowed ({100.0 * no.value / max(ns.value, 1):.2f} %)")
out = torch.empty(nq, dtype=torch.int32, device=dev)
qlen = int(os.environ.get("QLEN", 150))
only = os.environ.get("ONLY", "")
cases = []
if only in ("", "random"):
    cases.append(("random", synth.queries_fixed_len(43, nq, synth.GRCH38, qlen, device=dev)))
if only in ("", "sorted"):
    cases.append(("sorted", synth.queries_sorted_fixed_len(45, nq, synth.GRCH38, qlen, device=dev)))
stats = (C.c_uint64 * 2)()
for name, (qs, qe) in cases:
    for _ in range(int(os.environ.get("WARM", 2))):
        _ffi.check(lib.lapper_count_batch_u32_dev(h, C.c_void_p(qs.data_ptr()), C.c_void_p(qe.data_ptr()), nq,
                                                  C.c_void_p(out.data_ptr()), sp))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        _ffi.check(lib.lapper_count_batch_u32_dev(h, C.c_void_p(qs.data_ptr()), C.c_void_p(qe.data_ptr()), nq,
                                                  C.c_void_p(out.data_ptr()), sp))
    b.record()
    torch.cuda.synchronize()
    lib.lapper_debug_sorted_run_stats(stats, 1)
    print(f"{name}: {a.elapsed_time(b) / reps:.3f} ms per launch, checksum {int(out.to(torch.int64).sum())}; "
          f"sorted-run kernel: {stats[0]} runs ranked, {stats[1]} per query")
lib.lapper_free(h)
