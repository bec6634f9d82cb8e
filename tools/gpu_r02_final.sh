#!/bin/bash
# Round 2, final state: full GPU suite, smoke, bench (both arms), step breakdown, ncu launch list of the bench step and
# `--set full` captures of the dominant kernel and the new phase-refinement kernel.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/r02_final_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02_final_pytest.log; tail -4 $OUT/r02_final_pytest.log
timeout -s KILL 300 python __graft_entry__.py smoke > $OUT/r02_final_smoke.log 2>&1; tail -2 $OUT/r02_final_smoke.log
timeout -s KILL 900 python bench.py > $OUT/r02_final_bench.log 2>&1; tail -1 $OUT/r02_final_bench.log | cut -c1-400
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02_final_bench_reference.log 2>&1; tail -1 $OUT/r02_final_bench_reference.log | cut -c1-200
CPN_PROFILE_RANGE=step timeout -s KILL 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > $OUT/r02_final_ncu_bench.log 2>&1; tail -1 $OUT/r02_final_ncu_bench.log | cut -c1-120
timeout -s KILL 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -o $OUT/prof_final python tools/run_heads_op.py heads.block.0,core.refinement_head.block.0 fp16f8 > $OUT/r02_final_ncu_convs.log 2>&1; tail -2 $OUT/r02_final_ncu_convs.log
timeout -s KILL 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -o $OUT/prof_final_c2 python tools/run_heads_op.py core.refinement_head.block.0 fp16f8 CpnResNet18FPN 32 > $OUT/r02_final_ncu_c2.log 2>&1; tail -1 $OUT/r02_final_ncu_c2.log
find $OUT -size +45M -delete
