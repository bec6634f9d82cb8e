#!/bin/bash
# One-call GPU validation: parity tests, smoke, bench (both arms), micro-benches, per-op profiles, C4 slide with label
# rasterisation, ncu launch list + `--set full` captures of one kernel per class.  Everything lands in gpurun_out/;
# `python tools/make_profiles.py rNN` digests it into profiles/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -6 $OUT/pytest.log
timeout -s KILL 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout -s KILL 900 python bench.py > $OUT/bench_fp16.log 2>&1; tail -1 $OUT/bench_fp16.log | cut -c1-300
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.log 2>&1; tail -1 $OUT/bench_reference.log | cut -c1-300
timeout -s KILL 300 python tools/bench_decode.py > $OUT/bench_decode.log 2>&1; cat $OUT/bench_decode.log
timeout -s KILL 600 python tests/bench_c2l.py > $OUT/bench_c2l.log 2>&1; tail -1 $OUT/bench_c2l.log | cut -c1-300
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile.txt 2>&1; head -4 $OUT/plan_profile.txt
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16x3 > $OUT/plan_profile_fp16x3.txt 2>&1; head -2 $OUT/plan_profile_fp16x3.txt
timeout -s KILL 600 python tools/profile_plan.py CpnResNet18FPN 32 512 fp16 > $OUT/plan_profile_c2.txt 2>&1; head -2 $OUT/plan_profile_c2.txt
timeout -s KILL 300 python tools/profile_post.py > $OUT/profile_post.txt 2>&1; tail -16 $OUT/profile_post.txt
timeout -s KILL 900 python tools/run_wsi.py --size 16384 --labels > $OUT/wsi_16384_n1.log 2>&1; tail -1 $OUT/wsi_16384_n1.log
CPN_PROFILE_RANGE=step timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
OPS=heads.block.0,core.refinement_head.block.0,core.backbone.unet.layer_blocks.2.0,core.backbone.unet.layer_blocks.0.3,core.backbone.body.3.8.conv3,core.backbone.body.3.8.conv1,core.backbone.body.3.8.conv2,core.backbone.body.1.1.1.conv3,core.backbone.unet.inner_blocks.1
timeout -s KILL 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -o $OUT/prof_convs python tools/run_heads_op.py $OPS > $OUT/ncu_convs.log 2>&1; tail -2 $OUT/ncu_convs.log
timeout -s KILL 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'select_|decode_refine|nms_|gather_rows|prep_im2col|upsample|maxpool' -c 40 \
   -o $OUT/prof_post python tools/run_heads_op.py post > $OUT/ncu_post.log 2>&1; tail -2 $OUT/ncu_post.log
find $OUT -size +45M -delete
ls -la $OUT | head -50
