#!/bin/bash
# Round 2, call v: CpnResUNet (ResBlock U-Net; embedded 1x1 on the im2col stem) against the oracle, all engines.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -k "ragged and (ResUNet or WideU22 or CpnU22)" > $OUT/r02v_pytest_resunet.log 2>&1
echo "pytest rc=$?" >> $OUT/r02v_pytest_resunet.log; tail -25 $OUT/r02v_pytest_resunet.log
