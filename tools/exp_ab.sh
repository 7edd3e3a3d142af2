#!/bin/bash
# A/B of library variants (make VARIANT=x) on the cfg-2 count shape
for v in "" ${VARIANTS:-_pipe}; do
  lib=$PWD/rust_lapper_b200/liblapper_b200$v.so
  [ -f $lib ] || continue
  echo "== variant '$v'"; LAPPER_B200_LIB=$lib REPS=${REPS:-20} WARM=3 python tools/prof_count.py 2>&1 | tail -2
done
