import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rust_lapper_b200 import Lapper
from oracle import lapper_oracle as oracle
X = 0xFFFFFFFF
s = np.array([0, 0, X - 5, X, X - 1, 7, X], dtype=np.uint32)
e = np.array([X, 0, X, X, X, 7, X - 1], dtype=np.uint32)
vals = [0, 1, 5, 7, X - 6, X - 5, X - 1, X]
qs, qe = np.meshgrid(vals, vals)
qs, qe = qs.ravel().astype(np.uint32), qe.ravel().astype(np.uint32)
qs, qe = np.tile(qs, 40), np.tile(qe, 40)
cpu = oracle.OracleLapper(starts=s, stops=e)
want = cpu.count_batch(qs, qe)
for flags in (0, 1, 2, 2 | ((13 + 1) << 8), 2 | 8, 2 | (1820 << 16), 2 | (3 << 16)):
    gpu = Lapper(starts=s, stops=e, flags=flags)
    got = gpu.count_batch(qs, qe)
    bad = np.nonzero(got != want)[0]
    print("flags", hex(flags), "shift", gpu.bucket_shift, "packed", gpu.packed_table, "mismatches", bad.size)
    for j in bad[:12]:
        print("   q", j, int(qs[j]), int(qe[j]), "got", int(got[j]), "want", int(want[j]))
    got32 = gpu.count_batch_c32(qs, qe)
    bad = np.nonzero(got32 != want.astype(np.uint32))[0]
    print("   c32 mismatches", bad.size, [(int(qs[j]), int(qe[j]), int(got32[j]), int(want[j])) for j in bad[:6]])
