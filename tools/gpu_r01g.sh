#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -x -k "input_contract or tiled_im2col or batch_invariance or strict_fp32 or fp16_tensor" -p no:cacheprovider > $OUT/pytest_g.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_g.log; tail -8 $OUT/pytest_g.log
timeout -s KILL 600 python bench.py --no-cpu-baseline > $OUT/bench_tiledprep.log 2>&1; tail -1 $OUT/bench_tiledprep.log | cut -c1-300
CPN_PREP_TILED=0 timeout -s KILL 600 python bench.py --no-cpu-baseline > $OUT/bench_directprep.log 2>&1; tail -1 $OUT/bench_directprep.log | cut -c1-300
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile.txt 2>&1; grep "prep" $OUT/plan_profile.txt | head -3
