#!/bin/bash
# Round 2, call f: CTA-pair kernel (cta_group::2) for the 1x1 layers -- parity first (bail out on the first hang), then A/B.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
for e in tcgen05 tcgen05f8 tcgen05x3; do
  for c in 11 5 12; do
    timeout -s KILL 50 python tests/gpu_conv_check.py $e $c > $OUT/r02f_conv_${e}_$c.log 2>&1; rc=$?
    cut -c1-170 $OUT/r02f_conv_${e}_$c.log | tail -2
    if [ $rc -ne 0 ]; then echo "FAILED $e $c rc=$rc -- stopping"; exit 0; fi
  done
done
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -x -k "gate_passing or conv_tcgen05 or batch_invariance or ragged_input_sizes" > $OUT/r02f_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $OUT/r02f_pytest.log; tail -8 $OUT/r02f_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
OPS=core.backbone.body.3.8.conv1,core.backbone.body.3.8.conv3,core.backbone.body.1.1.1.conv1,core.backbone.body.1.1.1.conv3,core.backbone.body.2.1.conv3,core.backbone.unet.inner_blocks.1,core.backbone.body.4.1.conv3,core.backbone.body.2.0.conv1
: > $OUT/r02f_ab.log
env CPN_PAIR=0 timeout -s KILL 120 python tools/profile_ops.py fp16f8 $OPS >> $OUT/r02f_ab.log 2>&1
env CPN_PAIR=1 timeout -s KILL 120 python tools/profile_ops.py fp16f8 $OPS >> $OUT/r02f_ab.log 2>&1
env CPN_PAIR=0 timeout -s KILL 120 python tools/profile_ops.py fp16 $OPS >> $OUT/r02f_ab.log 2>&1
env CPN_PAIR=1 timeout -s KILL 120 python tools/profile_ops.py fp16 $OPS >> $OUT/r02f_ab.log 2>&1
cut -c1-160 $OUT/r02f_ab.log
