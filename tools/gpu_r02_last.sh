#!/bin/bash
# Round 2, last sanity of HEAD: the tiled driver (plain, repetitions, preprocessing chain) after the host-side refactors.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 150 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 120 -p no:cacheprovider -k "apply_model_matches or repetitions or preprocesses" > $OUT/r02_last_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02_last_pytest.log; tail -5 $OUT/r02_last_pytest.log
