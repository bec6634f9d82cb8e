#!/bin/bash
# Round 2, call g: why is the CTA-pair kernel slower per op?  ncu capture + switches.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
OPS=core.backbone.body.3.8.conv1,core.backbone.body.3.8.conv3,core.backbone.body.1.1.1.conv3
: > $OUT/r02g_ab.log
env CPN_PAIR=1 timeout -s KILL 120 python tools/profile_ops.py fp16f8 $OPS >> $OUT/r02g_ab.log 2>&1
env CPN_PAIR=1 CPN_DBG_EPI=1 timeout -s KILL 120 python tools/profile_ops.py fp16f8 $OPS >> $OUT/r02g_ab.log 2>&1
env CPN_PAIR=1 CPN_TC_STAGES=3 timeout -s KILL 120 python tools/profile_ops.py fp16f8 $OPS >> $OUT/r02g_ab.log 2>&1
env CPN_PAIR=1 CPN_TC_STAGES=4 timeout -s KILL 120 python tools/profile_ops.py fp16f8 $OPS >> $OUT/r02g_ab.log 2>&1
env CPN_PAIR=0 CPN_DBG_EPI=1 timeout -s KILL 120 python tools/profile_ops.py fp16f8 $OPS >> $OUT/r02g_ab.log 2>&1
cut -c1-160 $OUT/r02g_ab.log
CPN_PAIR=1 timeout -s KILL 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -o $OUT/prof_pair python tools/run_heads_op.py $OPS fp16f8 > $OUT/r02g_ncu.log 2>&1; tail -2 $OUT/r02g_ncu.log
