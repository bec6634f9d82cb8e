#!/bin/bash
# Round 2, call o: final state -- full GPU suite, smoke, bench (both arms), step breakdown.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/r02o_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02o_pytest.log; tail -4 $OUT/r02o_pytest.log
timeout -s KILL 300 python __graft_entry__.py smoke > $OUT/r02o_smoke.log 2>&1; tail -3 $OUT/r02o_smoke.log
timeout -s KILL 900 python bench.py > $OUT/r02o_bench.log 2>&1; tail -1 $OUT/r02o_bench.log | cut -c1-600
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02o_bench_reference.log 2>&1; tail -1 $OUT/r02o_bench_reference.log | cut -c1-200
timeout -s KILL 600 python tools/profile_step.py > $OUT/r02o_profile_step.log 2>&1; tail -3 $OUT/r02o_profile_step.log
