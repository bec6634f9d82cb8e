#!/bin/bash
# Round 2, call k: full GPU suite on the final code, ncu launch list + per-class captures (incl. the sparse-heads kernels).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --durations=8 > $OUT/r02k_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02k_pytest.log; tail -16 $OUT/r02k_pytest.log
CPN_PROFILE_RANGE=step timeout -s KILL 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > $OUT/r02k_ncu_bench.log 2>&1; tail -1 $OUT/r02k_ncu_bench.log | cut -c1-150
OPS=heads.block.0,core.refinement_head.block.0,core.backbone.unet.layer_blocks.2.0,core.backbone.unet.layer_blocks.0.0,core.backbone.unet.layer_blocks.0.3,core.backbone.body.3.8.conv3,core.backbone.body.3.8.conv1,core.backbone.body.3.8.conv2,core.backbone.body.1.1.1.conv3,core.backbone.unet.inner_blocks.1
timeout -s KILL 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -o $OUT/prof_convs_f8 python tools/run_heads_op.py $OPS fp16f8 > $OUT/r02k_ncu_convs.log 2>&1; tail -2 $OUT/r02k_ncu_convs.log
timeout -s KILL 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'gather_patches|conv_tc_kernel|select_|decode_refine|nms_|gather_rows|prep_im2col|upsample|maxpool' -c 40 \
   -o $OUT/prof_post python tools/run_heads_op.py post fp16f8 > $OUT/r02k_ncu_post.log 2>&1; tail -2 $OUT/r02k_ncu_post.log
find $OUT -size +45M -delete
