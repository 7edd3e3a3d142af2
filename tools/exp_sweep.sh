#!/bin/bash
# experiment: count kernel variants x table sizes (device-resident, random + sorted queries)
for v in 0 1 2 3 4 5 7; do
  for n in 10000000; do
    echo "== variant $v n=$n"; LAPPER_COUNT_VARIANT=$v N_IV=$n REPS=5 python tools/prof_count.py 2>&1 | tail -3
  done
done
for n in 1000000 2500000 5000000 7500000; do
  echo "== variant 0 n=$n"; LAPPER_COUNT_VARIANT=0 N_IV=$n REPS=5 python tools/prof_count.py 2>&1 | tail -2
  echo "== variant 3 n=$n"; LAPPER_COUNT_VARIANT=3 N_IV=$n REPS=5 python tools/prof_count.py 2>&1 | tail -2
done
