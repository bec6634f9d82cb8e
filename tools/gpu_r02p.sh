#!/bin/bash
# Round 2, call p: decoder convs over cat(lateral, up(top)) as phase conv + lateral conv -- parity, A/B.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x -k "gate_passing or batch_invariance or variant or sparse or ragged or full_size_c3" > $OUT/r02p_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $OUT/r02p_pytest.log; tail -12 $OUT/r02p_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
env CPN_UP2=0 timeout -s KILL 600 python bench.py --quick --no-cpu-baseline > $OUT/r02p_bench_noup2.log 2>&1; tail -1 $OUT/r02p_bench_noup2.log | cut -c1-230
timeout -s KILL 600 python bench.py --quick --no-cpu-baseline > $OUT/r02p_bench_up2.log 2>&1; tail -1 $OUT/r02p_bench_up2.log | cut -c1-230
env CPN_UP2_SKIP=0 timeout -s KILL 600 python bench.py --quick --no-cpu-baseline > $OUT/r02p_bench_up2_noskip.log 2>&1; tail -1 $OUT/r02p_bench_up2_noskip.log | cut -c1-230
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16f8 > $OUT/plan_profile_fp16f8_p.txt 2>&1; head -22 $OUT/plan_profile_fp16f8_p.txt; tail -13 $OUT/plan_profile_fp16f8_p.txt
