#!/bin/bash
# Round 2, call h: tap-pair kernel for the 64-wide layers -- parity first (bail out on the first failure), then A/B.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
for e in tcgen05f8 tcgen05 tcgen05x3; do
  for c in 1 3 7 9 10; do
    timeout -s KILL 50 python tests/gpu_conv_check.py $e $c > $OUT/r02h_conv_${e}_$c.log 2>&1; rc=$?
    cut -c1-170 $OUT/r02h_conv_${e}_$c.log | tail -2
    if [ $rc -ne 0 ]; then echo "FAILED $e $c rc=$rc -- stopping"; exit 0; fi
  done
done
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 180 -p no:cacheprovider -x -k "gate_passing or conv_tcgen05 or batch_invariance or ragged_input_sizes or variant or cuda_graph or input_contract" > $OUT/r02h_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $OUT/r02h_pytest.log; tail -8 $OUT/r02h_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
OPS=core.refinement_head.block.0,core.backbone.unet.layer_blocks.0.0,core.backbone.unet.layer_blocks.0.3,core.backbone.body.3.8.conv2,core.backbone.body.1.1.1.conv2,core.backbone.body.2.1.conv2,core.backbone.body.4.1.conv2
: > $OUT/r02h_ab.log
env CPN_TAP2=0 timeout -s KILL 120 python tools/profile_ops.py fp16f8 $OPS >> $OUT/r02h_ab.log 2>&1
env CPN_TAP2=1 timeout -s KILL 120 python tools/profile_ops.py fp16f8 $OPS >> $OUT/r02h_ab.log 2>&1
env CPN_TAP2=0 timeout -s KILL 120 python tools/profile_ops.py fp16 $OPS >> $OUT/r02h_ab.log 2>&1
env CPN_TAP2=1 timeout -s KILL 120 python tools/profile_ops.py fp16 $OPS >> $OUT/r02h_ab.log 2>&1
cut -c1-160 $OUT/r02h_ab.log
