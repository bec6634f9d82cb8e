#!/bin/bash
# Coalesced-epilogue A/B (same box): conv parity tests, model tests, bench + per-op profile with the line-coalesced epilogue
# and with it switched off in the halo kernel (CPN_COALESCE_HALO=0; CPN_COALESCE=0 switches it off everywhere).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -q --timeout 600 -x -p no:cacheprovider > $OUT/pytest_d.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_d.log; tail -12 $OUT/pytest_d.log
timeout -s KILL 600 python bench.py --no-cpu-baseline > $OUT/bench_coal.log 2>&1; tail -1 $OUT/bench_coal.log | cut -c1-330
CPN_COALESCE_HALO=0 timeout -s KILL 600 python bench.py --no-cpu-baseline > $OUT/bench_nocoalhalo.log 2>&1; tail -1 $OUT/bench_nocoalhalo.log | cut -c1-330
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile.txt 2>&1; grep "by class" -A 10 $OUT/plan_profile.txt; grep "k1 s1" $OUT/plan_profile.txt | head -8
timeout -s KILL 600 python tools/profile_plan.py CpnResNet18FPN 32 512 fp16 > $OUT/plan_profile_c2.txt 2>&1; head -3 $OUT/plan_profile_c2.txt
