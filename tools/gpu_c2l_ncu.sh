#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 170 ncu --set full --clock-control none -k regex:'c2l_|rlc_' -c 2 -o $OUT/prof_c2l python tests/bench_c2l.py > $OUT/ncu_c2l.log 2>&1; tail -2 $OUT/ncu_c2l.log | cut -c1-200
