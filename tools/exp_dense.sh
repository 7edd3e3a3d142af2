#!/bin/bash
# packed-table A/B on the cfg-2 shape: without it, with it at several fill targets
echo "== no packed table"; FLAGS=4 REPS=10 python tools/prof_count.py 2>&1 | tail -3
for lam in ${LAMS:-4.0 4.5 5.0 5.5 6.0}; do
  echo "== packed, starts/sector $lam"; LAPPER_DENSE_STARTS_PER_SECTOR=$lam REPS=10 python tools/prof_count.py 2>&1 | tail -3
done
