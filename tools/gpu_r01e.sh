#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_labels.py tests/test_gpu_model.py -m gpu -q --timeout 600 -k "label or apply_model" -p no:cacheprovider > $OUT/pytest_e.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_e.log; tail -12 $OUT/pytest_e.log
timeout -s KILL 300 python tools/run_wsi.py --size 8192 > $OUT/det3_n1.log 2>&1; tail -1 $OUT/det3_n1.log
