#!/bin/bash
for e in 0 1 2 3 4 6 7; do echo "== LAPPER_EXP=$e"; LAPPER_EXP=$e REPS=5 python tools/prof_count.py 2>&1 | tail -2; done
echo "== n=1M"; for e in 0 1 3 7; do echo "== LAPPER_EXP=$e"; N_IV=1000000 LAPPER_EXP=$e REPS=5 python tools/prof_count.py 2>&1 | tail -2; done
