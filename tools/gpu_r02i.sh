#!/bin/bash
# Round 2, call i: phase-decomposed up-sample + 3x3 (CPN_CONV_UP2) and the tap-pair rule -- parity, A/B, bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 180 -p no:cacheprovider -x -k "gate_passing or batch_invariance or ragged_input_sizes or variant or cuda_graph or full_size_c3" > $OUT/r02i_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $OUT/r02i_pytest.log; tail -8 $OUT/r02i_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
OPS=core.refinement_head.block.0,core.backbone.unet.layer_blocks.0.0,core.backbone.unet.layer_blocks.0.3
: > $OUT/r02i_ab.log
env CPN_UP2=0 timeout -s KILL 120 python tools/profile_ops.py fp16f8 $OPS >> $OUT/r02i_ab.log 2>&1
env CPN_UP2=1 timeout -s KILL 120 python tools/profile_ops.py fp16f8 $OPS >> $OUT/r02i_ab.log 2>&1
env CPN_UP2=1 timeout -s KILL 120 python tools/profile_ops.py fp16 $OPS >> $OUT/r02i_ab.log 2>&1
cut -c1-160 $OUT/r02i_ab.log
timeout -s KILL 600 python bench.py --no-cpu-baseline > $OUT/r02i_bench.log 2>&1; tail -1 $OUT/r02i_bench.log | cut -c1-900
