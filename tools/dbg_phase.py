import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import lapper_oracle
from rust_lapper_b200 import Lapper, synth
U = 40_000_000
s, e = synth.intervals_gene_exon(44, 300_000, U)
gpu, cpu = Lapper(starts=s, stops=e, flags=2), lapper_oracle.OracleLapper(starts=s, stops=e)
print('shift', gpu.bucket_shift)
nq = 300_000 + 77
sq = synth.queries_sorted_fixed_len(45, nq, U, 150)
g = gpu.count_batch_c32(*sq); c = cpu.count_batch(sq[0], sq[1], 4).astype(np.uint32)
bad = np.nonzero(g != c)[0]
print('FORCE', os.environ.get('LAPPER_FORCE_PHASES'), 'mismatches', bad.size, bad[:10], g[bad[:10]], c[bad[:10]])
if bad.size:
    print('bad index mod 1024', (bad % 1024)[:20], 'tiles', np.unique(bad // 1024)[:10], 'warps', np.unique((bad % 1024)//128)[:10])
