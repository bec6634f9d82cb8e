#!/bin/bash
# Round 2, call e: what bounds the 1x1 / grouped layers of the encoder in the 2-pass engine?  A/B over experiment switches.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out/r02e_ab.log; : > $OUT
OPS=core.backbone.body.3.8.conv1,core.backbone.body.3.8.conv2,core.backbone.body.3.8.conv3,core.backbone.body.1.1.1.conv1,core.backbone.body.1.1.1.conv2,core.backbone.body.1.1.1.conv3,core.backbone.body.2.1.conv3,core.backbone.unet.inner_blocks.1,core.backbone.body.4.1.conv3
run() { env "$@" timeout -s KILL 300 python tools/profile_ops.py fp16f8 $OPS >> $OUT 2>&1; }
run CPN_X=base
run CPN_BN_1X1=128
run CPN_BN_1X1=64
run CPN_TC_STAGES=2
run CPN_TC_STAGES=3
run CPN_COALESCE=0
run CPN_DBG_EPI=1
run CPN_DBG_EPI=1 CPN_BN_1X1=128
grep -v "^$" $OUT | cut -c1-200
