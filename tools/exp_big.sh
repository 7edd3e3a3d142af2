#!/bin/bash
# large interval sets: sector table alone (FLAGS=4) vs with the packed table (auto)
for n in ${NS:-10000000 25000000 50000000}; do for fl in 4 0; do echo "== n=$n flags=$fl"; FLAGS=$fl N_IV=$n REPS=5 python tools/prof_count.py 2>&1 | tail -3; done; done
