#!/usr/bin/env python
"""Generates tests/golden/*.npz: seeded synthetic inputs (rust_lapper_b200.synth) and the outputs of the CPU
oracle (oracle/lapper_oracle.c, itself pinned to the reference's known-answer tests by
tests/test_oracle_golden.py) for count / find / merge_overlaps / cov / union_and_intersect / depth.

The reference is a Rust crate and no Rust toolchain exists in this image, so these vectors come from the
pinned restatement, not from rustc output.  Re-run with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import lapper_oracle  # noqa: E402
from rust_lapper_b200 import synth  # noqa: E402

CASES = {
    # name: (interval generator, query generator)
    "cfg2_uniform": (lambda: synth.intervals_uniform(42, 3000, 4_000_000, 50, 5000),
                     lambda: synth.queries_fixed_len(43, 4000, 4_000_000, 150)),
    "cfg3_gene_exon_sorted": (lambda: synth.intervals_gene_exon(44, 3000, 6_000_000),
                              lambda: synth.queries_sorted_fixed_len(45, 4000, 6_000_000, 150)),
    "cfg4_pathological": (lambda: synth.intervals_pathological(46, 2000, 50_000_000),
                          lambda: synth.queries_fixed_len(47, 300, 50_000_000, 150)),
    "cfg1a_bakeoff": (lambda: synth.bakeoff_sets(1, 2, 400, 100_000)[0],
                      lambda: synth.bakeoff_sets(1, 2, 400, 100_000)[1]),
}


def main():
    for name, (gen_iv, gen_q) in CASES.items():
        s, e = gen_iv()
        qs, qe = gen_q()
        lap = lapper_oracle.OracleLapper(starts=s, stops=e)
        counts = lap.count_batch(qs, qe)
        off, idx = lap.find_batch(qs, qe)
        ss, se, sv = lap.intervals_soa()
        other = lapper_oracle.OracleLapper(starts=qs, stops=qe)
        ui = lap.union_and_intersect(other)
        cov = lap.cov()
        depth = np.array(lap.depth(), dtype=np.uint32).reshape(-1, 3) if name != "cfg4_pathological" else np.zeros((0, 3), np.uint32)
        lap.merge_overlaps()
        ms, me, mv = lap.intervals_soa()
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), starts=s, stops=e, q_start=qs, q_stop=qe, counts=counts, offsets=off,
            indices=idx, sorted_starts=ss, sorted_stops=se, perm=sv, max_len=np.uint32(lapper_oracle.OracleLapper(starts=s, stops=e).max_len),
            cov=np.uint32(cov), union_intersect_vs_queries=np.array(ui, np.uint32), depth=depth, merged_starts=ms,
            merged_stops=me, merged_perm=mv)
        print(name, "intervals", s.size, "queries", qs.size, "hits", int(off[-1]), "depth runs", depth.shape[0])


if __name__ == "__main__":
    main()
