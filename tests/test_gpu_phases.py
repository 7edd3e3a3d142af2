"""The count kernel's table phases (csrc/query.cu PhaseCtl) only switch on for tables larger than L2 can
hold; this forces 2..4 phases on small inputs (in a subprocess: the hook is read once per process) and
checks counts and find's CSR against the oracle, for random, sorted and mixed query orders."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from oracle import lapper_oracle
from rust_lapper_b200 import Lapper, synth
U = 40_000_000
s, e = synth.intervals_gene_exon(44, 300_000, U)
gpu, cpu = Lapper(starts=s, stops=e, flags=2), lapper_oracle.OracleLapper(starts=s, stops=e)
nq = 300_000 + 77
rq = synth.queries_fixed_len(43, nq, U, 150)
sq = synth.queries_sorted_fixed_len(45, nq, U, 150)
mixed = (np.concatenate([sq[0][:100_000], rq[0][:100_077], sq[0][200_000:300_000]]),
         np.concatenate([sq[1][:100_000], rq[1][:100_077], sq[1][200_000:300_000]]))
long_q = (rq[0], np.minimum(rq[0].astype(np.int64) + 30_000, 0xFFFFFFFF).astype(np.uint32))
inverted = (rq[1], rq[0])
for name, (qs, qe) in (("random", rq), ("sorted", sq), ("mixed", mixed), ("long", long_q), ("inverted", inverted)):
    assert np.array_equal(gpu.count_batch_c32(qs, qe), cpu.count_batch(qs, qe, 4).astype(np.uint32)), name
    go, gi = gpu.find_batch(qs, qe)
    co, ci = cpu.find_batch(qs, qe, 4)
    assert np.array_equal(go, co) and np.array_equal(gi, ci), name
print("phases ok")
"""


@pytest.mark.parametrize("phases", [2, 3, 4])
def test_forced_phases_match_oracle(phases):
    env = dict(os.environ, LAPPER_FORCE_PHASES=str(phases))
    res = subprocess.run([sys.executable, "-c", SCRIPT % (ROOT, os.path.join(ROOT, "tests"))], env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "phases ok" in res.stdout, res.stdout + res.stderr
