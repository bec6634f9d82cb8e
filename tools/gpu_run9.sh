#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_post.py tests/test_gpu_model.py -m gpu -q --timeout 900 -p no:cacheprovider -k "stitching or apply_model or border" > $OUT/pytest_small.log 2>&1; tail -3 $OUT/pytest_small.log
timeout -s KILL 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -o $OUT/prof_ref_head python tools/run_heads_op.py core.refinement_head.block.0 > $OUT/ncu_ref_head.log 2>&1; tail -2 $OUT/ncu_ref_head.log
timeout -s KILL 900 ncu --profile-from-start off --set full --clock-control none \
   -o $OUT/prof_1x1 python tools/run_heads_op.py core.backbone.body.3.5.conv3 > $OUT/ncu_1x1.log 2>&1; tail -2 $OUT/ncu_1x1.log
find $OUT -size +40M -delete
