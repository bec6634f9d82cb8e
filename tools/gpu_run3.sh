#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -15 $OUT/pytest.log
CPN_F2C_MINB=3 timeout -s KILL 300 python tools/bench_decode.py > $OUT/bench_decode_minb3.log 2>&1; cat $OUT/bench_decode_minb3.log
CPN_F2C_MINB=4 timeout -s KILL 300 python tools/bench_decode.py > $OUT/bench_decode_minb4.log 2>&1; cat $OUT/bench_decode_minb4.log
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile.txt 2>&1; head -40 $OUT/plan_profile.txt; tail -14 $OUT/plan_profile.txt
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_fp16.log 2>&1; tail -2 $OUT/bench_fp16.log
# ncu launch list of the timed region only (profiler range), then a full capture of every conv_tc launch of one step
CPN_PROFILE_RANGE=step timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
CPN_PROFILE_RANGE=step timeout -s KILL 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:conv_tc_kernel -c 140 -o $OUT/prof_conv_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT | head -40
