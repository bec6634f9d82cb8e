#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 300 python tests/gpu_conv_check.py tcgen05 0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 > $OUT/conv_checks.log 2>&1; cut -c1-150 $OUT/conv_checks.log
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -8 $OUT/pytest.log
echo "== rotate=1"; timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile.txt 2>&1; head -20 $OUT/plan_profile.txt; tail -11 $OUT/plan_profile.txt
echo "== rotate=0"; CPN_ROTATE=0 timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile_norot.txt 2>&1; head -8 $OUT/plan_profile_norot.txt; tail -11 $OUT/plan_profile_norot.txt
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_fp16.log 2>&1; tail -1 $OUT/bench_fp16.log | cut -c1-300
CPN_ROTATE=0 timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_norot.log 2>&1; tail -1 $OUT/bench_norot.log | cut -c1-300
timeout -s KILL 600 python tools/layer_diff.py CpnResNeXt101UNet 256 > $OUT/layer_diff.txt 2>&1; tail -30 $OUT/layer_diff.txt
timeout -s KILL 600 python tools/profile_plan.py CpnResNet18FPN 32 512 fp16 > $OUT/plan_profile_c2.txt 2>&1; head -12 $OUT/plan_profile_c2.txt
timeout -s KILL 600 python tools/profile_plan.py CpnU22 16 256 fp16 > $OUT/plan_profile_c1.txt 2>&1; head -8 $OUT/plan_profile_c1.txt
find $OUT -size +40M -delete
