#!/bin/bash
# GPU call 2 of this session: new tests (ensembles, masks, contours2labels), bench after the residual-epilogue change,
# contours2labels micro-bench, C4 slide incl. label rasterisation, per-op profile.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -x -p no:cacheprovider > $OUT/pytest_new.log 2>&1
echo "pytest(new) rc=$?" >> $OUT/pytest_new.log; tail -25 $OUT/pytest_new.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -15 $OUT/pytest.log
timeout -s KILL 900 python bench.py > $OUT/bench_fp16.log 2>&1; tail -1 $OUT/bench_fp16.log | cut -c1-400
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile.txt 2>&1; head -4 $OUT/plan_profile.txt; grep "by class" -A 12 $OUT/plan_profile.txt
timeout -s KILL 600 python tools/bench_c2l.py > $OUT/bench_c2l.log 2>&1; tail -2 $OUT/bench_c2l.log
timeout -s KILL 900 python tools/run_wsi.py --size 16384 --labels > $OUT/wsi_16384_n1.log 2>&1; tail -1 $OUT/wsi_16384_n1.log
ls -la $OUT | head -40
