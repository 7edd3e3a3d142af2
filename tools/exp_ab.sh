#!/bin/bash
for v in "" _o4; do
  lib=$PWD/rust_lapper_b200/liblapper_b200$v.so
  [ -f $lib ] || continue
  for n in 10000000 1000000 50000000; do
  echo "== variant '$v' n=$n"; N_IV=$n LAPPER_B200_LIB=$lib REPS=10 python tools/prof_count.py 2>&1 | tail -2
  done
done
