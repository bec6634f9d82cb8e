#!/bin/bash
# Round 2, call a: first contact of the 2-pass engine (fp16 + e4m3 corrections) with the hardware.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout -s KILL 300 python tests/gpu_conv_check.py tcgen05f8 0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 > $OUT/r02a_conv_f8.log 2>&1; cat $OUT/r02a_conv_f8.log | cut -c1-200
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x -k "gate or fp16_plus or input_contract or batch_invariance or variant" > $OUT/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02a_pytest.log; tail -30 $OUT/r02a_pytest.log
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16f8 > $OUT/plan_profile_fp16f8.txt 2>&1; head -30 $OUT/plan_profile_fp16f8.txt; tail -14 $OUT/plan_profile_fp16f8.txt
timeout -s KILL 900 python bench.py --steps 5 --no-cpu-baseline > $OUT/r02a_bench.log 2>&1; tail -1 $OUT/r02a_bench.log | cut -c1-1500
