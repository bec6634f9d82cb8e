#!/bin/bash
# Round 2, call m: bias prefetch in the coalesced epilogue -- parity + timing.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 300 python tests/gpu_conv_check.py tcgen05f8 > $OUT/r02m_conv_f8.log 2>&1; cut -c1-150 $OUT/r02m_conv_f8.log | tail -17
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 180 -p no:cacheprovider -x -k "gate_passing or conv_tcgen05 or batch_invariance or variant or sparse" > $OUT/r02m_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $OUT/r02m_pytest.log; tail -4 $OUT/r02m_pytest.log
timeout -s KILL 600 python bench.py --quick --no-cpu-baseline > $OUT/r02m_bench.log 2>&1; tail -1 $OUT/r02m_bench.log | cut -c1-400
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16f8 > $OUT/plan_profile_fp16f8_m.txt 2>&1; tail -14 $OUT/plan_profile_fp16f8_m.txt
