#!/bin/bash
# Coalesced epilogue for the split-precision (fp16x3) engine: parity tests + same-box A/B.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -q --timeout 600 -x -k "conv or fp16x3 or ragged" -p no:cacheprovider > $OUT/pytest_f.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_f.log; tail -8 $OUT/pytest_f.log
timeout -s KILL 600 python bench.py --precision fp16x3 --no-cpu-baseline --steps 5 > $OUT/bench_x3_coal.log 2>&1; tail -1 $OUT/bench_x3_coal.log | cut -c1-300
CPN_COALESCE_SPLIT=0 timeout -s KILL 600 python bench.py --precision fp16x3 --no-cpu-baseline --steps 5 > $OUT/bench_x3_direct.log 2>&1; tail -1 $OUT/bench_x3_direct.log | cut -c1-300
