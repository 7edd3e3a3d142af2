#!/bin/bash
for n in 25000000 50000000; do for fl in 0 8; do echo "== n=$n flags=$fl"; FLAGS=$fl N_IV=$n REPS=5 python tools/prof_count.py 2>&1 | tail -3; done; done
