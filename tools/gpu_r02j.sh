#!/bin/bash
# Round 2, call j: sparse location / fourier heads -- parity first, then the bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -x -k "gather_patches or sparse_heads" > $OUT/r02j_pytest_sparse.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $OUT/r02j_pytest_sparse.log; tail -25 $OUT/r02j_pytest_sparse.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/r02j_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $OUT/r02j_pytest.log; tail -8 $OUT/r02j_pytest.log
timeout -s KILL 300 python __graft_entry__.py smoke > $OUT/r02j_smoke.log 2>&1; tail -3 $OUT/r02j_smoke.log
timeout -s KILL 900 python bench.py > $OUT/r02j_bench.log 2>&1; tail -1 $OUT/r02j_bench.log | cut -c1-1200
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16f8 > $OUT/plan_profile_fp16f8_sparse.txt 2>&1; head -16 $OUT/plan_profile_fp16f8_sparse.txt; tail -14 $OUT/plan_profile_fp16f8_sparse.txt
