"""e2e find (host queries -> host CSR, lapper_find_batch_u32_host) against what the link delivers for the same bytes.
Knobs are read by the library at first use: LAPPER_FIND_CHUNK_LOG2, LAPPER_FIND_SLOTS (one process per setting)."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rust_lapper_b200 import _ffi, synth  # noqa: E402

U = 3_100_000_000
n, nq = 10_000_000, int(os.environ.get("NQ", 100_000_000))
dev = torch.device("cuda", 0)
lib = _ffi.lib()
vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
s, e = synth.intervals_gene_exon(44, n, U, device=dev)
h = C.c_void_p(0)
_ffi.check(lib.lapper_build_u32_dev(vp(s), vp(e), n, 0, 0, sp, C.byref(h)))
qs, qe = synth.queries_sorted_fixed_len(45, nq, U, 150, device=dev)
hq_s = torch.empty(nq, dtype=torch.int32, pin_memory=True)
hq_e = torch.empty(nq, dtype=torch.int32, pin_memory=True)
hq_s.copy_(qs)
hq_e.copy_(qe)
h_off = torch.empty(nq + 1, dtype=torch.int64, pin_memory=True)
tot = C.c_uint64(0)
_ffi.check(lib.lapper_find_batch_u32_host(h, vp(hq_s), vp(hq_e), nq, vp(h_off), None, 0, C.byref(tot)))
hits = int(tot.value)
h_idx = torch.empty(hits, dtype=torch.int32, pin_memory=True)
d2h = (nq + 1) * 8 + hits * 4


def run():
    _ffi.check(lib.lapper_find_batch_u32_host(h, vp(hq_s), vp(hq_e), nq, vp(h_off), vp(h_idx), hits, C.byref(tot)))


for _ in range(2):
    run()
t0 = time.perf_counter()
K = 4
for _ in range(K):
    run()
dt = (time.perf_counter() - t0) / K
print(f"find_host: chunk 2^{os.environ.get('LAPPER_FIND_CHUNK_LOG2', '22')} slots {os.environ.get('LAPPER_FIND_SLOTS', '3')}: "
      f"{dt * 1e3:.2f} ms per {nq} queries ({hits} hits, {d2h / 1e9:.2f} GB out, {nq * 8 / 1e9:.2f} GB in) -> out {d2h / dt / 1e9:.1f} GB/s")

if os.environ.get("LINK", "0") == "1":
    # what the link gives for the same bytes: D2H of d2h bytes in 192 MB pieces, alone and with the 0.8 GB H2D beside it
    d_big = torch.empty(d2h // 4, dtype=torch.int32, device=dev)
    h_big = torch.empty(d2h // 4, dtype=torch.int32, pin_memory=True)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    piece = 48_000_000

    def d2h_all():
        with torch.cuda.stream(s1):
            for lo in range(0, d_big.numel(), piece):
                h_big[lo:lo + piece].copy_(d_big[lo:lo + piece], non_blocking=True)

    def h2d_all():
        with torch.cuda.stream(s2):
            qs.copy_(hq_s, non_blocking=True)
            qe.copy_(hq_e, non_blocking=True)

    for name, fns in (("D2H alone", (d2h_all,)), ("D2H + H2D of the queries", (d2h_all, h2d_all))):
        for _ in range(2):
            for f in fns:
                f()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            for f in fns:
                f()
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"link: {name}: {dt * 1e3:.2f} ms -> {d2h / dt / 1e9:.1f} GB/s out")
lib.lapper_free(h)
