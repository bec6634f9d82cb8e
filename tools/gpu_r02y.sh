#!/bin/bash
# Round 2, call y: vectorised fp16+e4m3 bilinear kernel -- FPN fixtures / ragged sizes / phase equality, C2 plain-path profile.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -k "phase_refinement or FPN or resnet18fpn or head_options or variant" > $OUT/r02y_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02y_pytest.log; tail -8 $OUT/r02y_pytest.log
CPN_REF_PHASE=0 timeout -s KILL 600 python tools/profile_plan.py CpnResNet18FPN 32 512 fp16f8 > $OUT/r02y_plan_profile_c2_plain.txt 2>&1
head -4 $OUT/r02y_plan_profile_c2_plain.txt; tail -1 $OUT/r02y_plan_profile_c2_plain.txt
CPN_REF_PHASE=1 timeout -s KILL 600 python tools/profile_plan.py CpnResNet18FPN 32 512 fp16f8 > $OUT/r02y_plan_profile_c2_phase.txt 2>&1
head -4 $OUT/r02y_plan_profile_c2_phase.txt; tail -1 $OUT/r02y_plan_profile_c2_phase.txt
