#!/bin/bash
# Round 2, call w: phase-decomposed refinement head (C2 family): parity fixtures, then C2 A/B.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "resnet18fpn or FPN or c2 or C2" > $OUT/r02w_pytest_fpn.log 2>&1
echo "pytest rc=$?" >> $OUT/r02w_pytest_fpn.log; tail -30 $OUT/r02w_pytest_fpn.log
