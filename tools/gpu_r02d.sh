#!/bin/bash
# Round 2, call d: evidence -- ncu launch list of the bench step, `--set full` captures per conv class in the 2-pass
# engine, compute-sanitizer memcheck / racecheck over the synchronisation-heavy kernels.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
CPN_PROFILE_RANGE=step timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > $OUT/r02d_ncu_bench.log 2>&1; tail -2 $OUT/r02d_ncu_bench.log | cut -c1-200
OPS=heads.block.0,core.refinement_head.block.0,core.backbone.unet.layer_blocks.2.0,core.backbone.unet.layer_blocks.0.3,core.backbone.body.3.8.conv3,core.backbone.body.3.8.conv1,core.backbone.body.3.8.conv2,core.backbone.body.1.1.1.conv3,core.backbone.unet.inner_blocks.1
timeout -s KILL 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -o $OUT/prof_convs_f8 python tools/run_heads_op.py $OPS fp16f8 > $OUT/r02d_ncu_convs.log 2>&1; tail -2 $OUT/r02d_ncu_convs.log
for tool in memcheck racecheck; do
  for c in conv labels nms; do
    timeout -s KILL 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py $c > $OUT/r02d_sanitizer_${tool}_${c}.log 2>&1
    echo "== $tool $c rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|hazard" $OUT/r02d_sanitizer_${tool}_${c}.log | head -8
  done
done
find $OUT -size +45M -delete
