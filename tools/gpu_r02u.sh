#!/bin/bash
# Round 2, call u: in-house radix sort behind every NMS variant -- post / labels / model suites, quick bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_gpu_post.py tests/test_gpu_labels.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/r02u_pytest_post.log 2>&1
echo "pytest rc=$?" >> $OUT/r02u_pytest_post.log; tail -6 $OUT/r02u_pytest_post.log
timeout -s KILL 600 python bench.py --quick --no-cpu-baseline --steps 20 > $OUT/r02u_bench.log 2>&1; tail -1 $OUT/r02u_bench.log | cut -c1-300
timeout -s KILL 600 python tools/profile_post.py > $OUT/r02u_profile_post.log 2>&1; tail -25 $OUT/r02u_profile_post.log
