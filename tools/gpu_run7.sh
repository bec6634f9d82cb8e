#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -6 $OUT/pytest.log
timeout -s KILL 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout -s KILL 900 python bench.py > $OUT/bench_fp16.log 2>&1; tail -1 $OUT/bench_fp16.log
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.log 2>&1; tail -1 $OUT/bench_reference.log | cut -c1-400
timeout -s KILL 600 python tools/profile_plan.py CpnResNet18FPN 32 512 fp16 > $OUT/plan_profile_c2.txt 2>&1; head -8 $OUT/plan_profile_c2.txt
timeout -s KILL 900 python tools/run_wsi.py --size 8192 > $OUT/wsi_8192_n1.log 2>&1; tail -1 $OUT/wsi_8192_n1.log
timeout -s KILL 900 python tools/run_wsi.py --size 16384 > $OUT/wsi_16384_n1.log 2>&1; tail -1 $OUT/wsi_16384_n1.log
find $OUT -size +40M -delete
