"""Profiling driver: cfg-3 find pipeline (gene/exon-like intervals x start-sorted 150-bp queries), device-resident."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rust_lapper_b200 import _ffi, synth  # noqa: E402

n = int(os.environ.get("N_IV", 10_000_000))
nq = int(os.environ.get("NQ", 100_000_000))
reps = int(os.environ.get("REPS", 3))
order = os.environ.get("ORDER", "sorted")
dev = torch.device("cuda", 0)
lib = _ffi.lib()
if os.environ.get("GEN", "gene_exon") == "uniform":  # the cfg-5 shape: GEN=uniform N_IV=50000000 NQ=20000000
    s, e = synth.intervals_uniform(42, n, synth.GRCH38, 50, 5000, device=dev)
else:
    s, e = synth.intervals_gene_exon(44, n, synth.GRCH38, device=dev)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
h = C.c_void_p(0)
_ffi.check(lib.lapper_build_u32_dev(C.c_void_p(s.data_ptr()), C.c_void_p(e.data_ptr()), n, 0, 0, sp, C.byref(h)))
if order == "sorted":
    qs, qe = synth.queries_sorted_fixed_len(45, nq, synth.GRCH38, 150, device=dev)
else:
    qs, qe = synth.queries_fixed_len(45, nq, synth.GRCH38, 150, device=dev)
off = torch.empty(nq + 1, dtype=torch.int64, device=dev)
total = C.c_uint64(0)
vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
_ffi.check(lib.lapper_find_batch_offsets_u32_dev(h, vp(qs), vp(qe), nq, vp(off), C.byref(total), sp))
idx = torch.empty(max(int(total.value), 1), dtype=torch.int32, device=dev)


if os.environ.get("ONE_CALL_ONLY"):
    # the one-call form alone (for ncu: every kernel of one lapper_find_batch_u32_dev call, after one warm-up call)
    off2, idx2, cap = torch.empty_like(off), torch.empty_like(idx), int(total.value)
    for _ in range(2):
        t = C.c_uint64(0)
        _ffi.check(lib.lapper_find_batch_u32_dev(h, vp(qs), vp(qe), nq, vp(off2), vp(idx2), cap, C.byref(t), sp))
    torch.cuda.synchronize()
    print(f"one call: {int(t.value)} hits, same CSR as the two-phase calls: {bool(torch.equal(off, off2))}")
    lib.lapper_free(h)
    sys.exit(0)


def step():
    t = C.c_uint64(0)
    _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h, vp(qs), vp(qe), nq, vp(off), C.byref(t), sp))
    _ffi.check(lib.lapper_find_batch_fill_u32_dev(h, vp(qs), vp(qe), nq, vp(off), vp(idx), sp))


for _ in range(int(os.environ.get("WARM", 1))):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
# phase split
torch.cuda.synchronize()
a.record()
for _ in range(reps):
    t = C.c_uint64(0)
    _ffi.check(lib.lapper_find_batch_offsets_u32_dev(h, vp(qs), vp(qe), nq, vp(off), C.byref(t), sp))
b.record()
torch.cuda.synchronize()
ms_off = a.elapsed_time(b) / reps
a.record()
for _ in range(reps):
    _ffi.check(lib.lapper_find_batch_fill_u32_dev(h, vp(qs), vp(qe), nq, vp(off), vp(idx), sp))
b.record()
torch.cuda.synchronize()
ms_fill = a.elapsed_time(b) / reps
stats = (C.c_uint64 * 2)()
lib.lapper_debug_sorted_run_stats(stats, 1)
print(f"  offsets phase {ms_off:.3f} ms, fill phase {ms_fill:.3f} ms; sorted-run kernel so far: {stats[0]} runs ranked, {stats[1]} per query")
print(f"find {order}: {ms:.3f} ms per pass, {int(total.value)} hits ({int(total.value) / nq:.2f}/query), "
      f"{nq / ms / 1e6:.2f} Gq/s, {int(total.value) / ms / 1e6:.1f} Ghits/s, checksum {int(idx[: int(total.value)].to(torch.int64).sum())}")
# the one-call form: count pass, scan and emit in one enqueue (the emit kernel derives the offsets)
off2 = torch.empty_like(off)
idx2 = torch.empty_like(idx)
cap = int(total.value)
for _ in range(2):
    t = C.c_uint64(0)
    _ffi.check(lib.lapper_find_batch_u32_dev(h, vp(qs), vp(qe), nq, vp(off2), vp(idx2), cap, C.byref(t), sp))
torch.cuda.synchronize()
a.record()
for _ in range(reps):
    t = C.c_uint64(0)
    _ffi.check(lib.lapper_find_batch_u32_dev(h, vp(qs), vp(qe), nq, vp(off2), vp(idx2), cap, C.byref(t), sp))
b.record()
torch.cuda.synchronize()
ms1 = a.elapsed_time(b) / reps
same = bool(torch.equal(off, off2)) and bool(torch.equal(idx[:cap], idx2[:cap]))
print(f"find {order}, one call: {ms1:.3f} ms per pass, {nq / ms1 / 1e6:.2f} Gq/s, {cap / ms1 / 1e6:.1f} Ghits/s, "
      f"same CSR as the two-phase calls: {same}")
if os.environ.get("EMIT_STATS"):
    import ctypes
    raw = ctypes.CDLL(_ffi.LIB_PATH)
    st = (ctypes.c_ulonglong * 8)()
    raw.lapper_debug_emit_stats(st, 1)
    _ffi.check(lib.lapper_find_batch_fill_u32_dev(h, vp(qs), vp(qe), nq, vp(off), vp(idx), sp))
    raw.lapper_debug_emit_stats(st, 0)
    print("  emit stats (tile kernel): nice tiles %d, fallback tiles %d, K/tile %.1f, raw/tile %.1f, fast groups %d (T %.1f), "
          "staged-slow groups %d, unstaged groups %d" % (st[0], st[2], st[1] / max(st[0], 1), st[3] / max(st[0], 1), st[4],
                                                         st[5] / max(st[4], 1), st[6], st[7]))
lib.lapper_free(h)
