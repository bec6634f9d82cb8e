#!/bin/bash
# Round 2, last call: CpnSlimU22 (zero-padded 32-channel layers) + the U-Net family and the flagship fixtures after the tracer change.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 500 -p no:cacheprovider -k "(ragged and (U22 or ResUNet or ResNeXt50UNet)) or gate_passing or strict_fp32" > $OUT/r02_slim_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02_slim_pytest.log; tail -12 $OUT/r02_slim_pytest.log
