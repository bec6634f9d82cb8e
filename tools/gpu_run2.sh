#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -15 $OUT/pytest.log
timeout -s KILL 300 python tools/bench_decode.py > $OUT/bench_decode.log 2>&1; cat $OUT/bench_decode.log
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile.txt 2>&1; head -70 $OUT/plan_profile.txt
timeout -s KILL 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
# ncu: launch list of the bench command (serialised, cold-cache: shares only) and a full capture of the dominant kernel
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 400 -c 3 -o $OUT/prof_conv_tc \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
