#!/bin/bash
for v in "" ${VARIANTS:-_e1 _e2 _e3}; do
  lib=$PWD/rust_lapper_b200/liblapper_b200$v.so
  [ -f $lib ] || continue
  echo "== variant '$v'"; LAPPER_B200_LIB=$lib REPS=5 python tools/prof_find.py 2>&1 | tail -2
done
