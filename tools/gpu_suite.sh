#!/bin/bash
# One-call GPU validation: parity tests, smoke, bench (both arms), per-op profile, ncu launch list + full capture of the
# dominant kernel.  Everything lands in gpurun_out/; `python tools/make_profiles.py rNN` digests it into profiles/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -6 $OUT/pytest.log
timeout -s KILL 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout -s KILL 900 python bench.py > $OUT/bench_fp16.log 2>&1; tail -1 $OUT/bench_fp16.log
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.log 2>&1; tail -1 $OUT/bench_reference.log | cut -c1-300
timeout -s KILL 300 python tools/bench_decode.py > $OUT/bench_decode.log 2>&1; cat $OUT/bench_decode.log
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile.txt 2>&1; head -8 $OUT/plan_profile.txt
timeout -s KILL 900 python tools/run_wsi.py --size 16384 > $OUT/wsi_16384_n1.log 2>&1; tail -1 $OUT/wsi_16384_n1.log
CPN_PROFILE_RANGE=step timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
timeout -s KILL 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -o $OUT/prof_heads_conv python tools/run_heads_op.py heads.block.0 > $OUT/ncu_full.log 2>&1
find $OUT -size +40M -delete
