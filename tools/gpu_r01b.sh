#!/bin/bash
# GPU call of the second session of round 1: parity suite incl. the new variants, A/B of the fp16x3 pass order, bench,
# ncu launch list, and `ncu --set full` captures of one kernel per class (conv shapes + post-head chain).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -15 $OUT/pytest.log
CPN_SPLIT_LOFIRST=0 CPN_PARITY_REPORT=$OUT/parity_report_hifirst.json timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -m gpu -q \
   -k "fp16x3" --timeout 600 -p no:cacheprovider > $OUT/pytest_hifirst.log 2>&1; tail -3 $OUT/pytest_hifirst.log
timeout -s KILL 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout -s KILL 900 python bench.py > $OUT/bench_fp16.log 2>&1; tail -1 $OUT/bench_fp16.log
CPN_PROFILE_RANGE=step timeout -s KILL 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
OPS=heads.block.0,core.refinement_head.block.0,core.backbone.unet.layer_blocks.2.0,core.backbone.unet.layer_blocks.0.3,core.backbone.body.3.8.conv3,core.backbone.body.3.8.conv1,core.backbone.body.3.8.conv2,core.backbone.body.1.1.1.conv3,core.backbone.unet.inner_blocks.1
timeout -s KILL 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -o $OUT/prof_convs python tools/run_heads_op.py $OPS > $OUT/ncu_convs.log 2>&1; tail -3 $OUT/ncu_convs.log
timeout -s KILL 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'select_|decode_refine|nms_|gather_rows|prep_im2col|upsample|maxpool' -c 40 \
   -o $OUT/prof_post python tools/run_heads_op.py post > $OUT/ncu_post.log 2>&1; tail -3 $OUT/ncu_post.log
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16x3 > $OUT/plan_profile_fp16x3.txt 2>&1; head -3 $OUT/plan_profile_fp16x3.txt
ls -la $OUT
find $OUT -size +45M -delete
