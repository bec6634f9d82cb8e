#!/bin/bash
# Round 2, call r: new GPU tests (slide preprocessing, test-time repetitions, output files) + micro-bench of the slide kernels.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_preprocess.py tests/test_outputs.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/r02r_pytest_new.log 2>&1
echo "pytest rc=$?" >> $OUT/r02r_pytest_new.log; tail -15 $OUT/r02r_pytest_new.log
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -k "repetitions or preprocesses or apply_model_matches" > $OUT/r02r_pytest_apply.log 2>&1
echo "pytest rc=$?" >> $OUT/r02r_pytest_apply.log; tail -15 $OUT/r02r_pytest_apply.log
timeout -s KILL 600 python tools/bench_preprocess.py > $OUT/r02r_bench_preprocess.log 2>&1; tail -20 $OUT/r02r_bench_preprocess.log
