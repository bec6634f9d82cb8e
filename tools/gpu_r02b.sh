#!/bin/bash
# Round 2, call b: full GPU suite with the new default engine, smoke, bench (N = 1) and the reference arm.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider --durations=12 > $OUT/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02b_pytest.log; tail -40 $OUT/r02b_pytest.log
timeout -s KILL 600 python __graft_entry__.py smoke > $OUT/r02b_smoke.log 2>&1; tail -3 $OUT/r02b_smoke.log
timeout -s KILL 900 python bench.py > $OUT/r02b_bench.log 2>&1; tail -1 $OUT/r02b_bench.log | cut -c1-3000
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02b_bench_reference.log 2>&1; tail -1 $OUT/r02b_bench_reference.log | cut -c1-400
