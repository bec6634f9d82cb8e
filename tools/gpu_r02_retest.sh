#!/bin/bash
# Round 2: the tiled-driver test whose pair criterion was made consistent with its count tolerance.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -k "tiled_driver_flagship" > $OUT/r02_final_pytest_retest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02_final_pytest_retest.log; tail -6 $OUT/r02_final_pytest_retest.log
