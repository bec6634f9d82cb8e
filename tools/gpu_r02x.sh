#!/bin/bash
# Round 2, call x: phase refinement head A/B -- equality test, C2 step time with and without (same box).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -k "phase_refinement" > $OUT/r02x_pytest_phase.log 2>&1
echo "pytest rc=$?" >> $OUT/r02x_pytest_phase.log; tail -25 $OUT/r02x_pytest_phase.log
for v in 0 1; do
  CPN_REF_PHASE=$v timeout -s KILL 600 python tools/profile_plan.py CpnResNet18FPN 32 512 fp16f8 > $OUT/r02x_plan_profile_c2_phase$v.txt 2>&1
  head -8 $OUT/r02x_plan_profile_c2_phase$v.txt; tail -1 $OUT/r02x_plan_profile_c2_phase$v.txt
done
