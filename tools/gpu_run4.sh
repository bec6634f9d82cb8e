#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
HALO_CASES="1 2 3 4 13 14 15"
check() { python - "$1" <<'PY'
import json,sys
rows=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')]
ok = rows and all(('error' not in r) and r['rel_err'] < 2e-3 for r in rows)
print('PASS' if ok else 'FAIL', [(r['case'], r.get('rel_err', r.get('error'))) for r in rows])
sys.exit(0 if ok else 1)
PY
}
timeout -s KILL 300 python tests/gpu_conv_check.py tcgen05 $HALO_CASES > $OUT/halo_default.log 2>&1
if check $OUT/halo_default.log; then echo "halo: default descriptor roles OK";
else
  CPN_HALO_SWAP=1 timeout -s KILL 300 python tests/gpu_conv_check.py tcgen05 $HALO_CASES > $OUT/halo_swap.log 2>&1
  if check $OUT/halo_swap.log; then echo "halo: SWAPPED roles OK"; export CPN_HALO_SWAP=1;
  else echo "halo: BROKEN -> disabled"; export CPN_HALO=0; CPN_DUMP=$OUT timeout -s KILL 120 python tests/gpu_conv_check.py tcgen05 1 >> $OUT/halo_default.log 2>&1; fi
fi
env | grep CPN_ > $OUT/halo_env.txt
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -15 $OUT/pytest.log
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile.txt 2>&1; head -24 $OUT/plan_profile.txt; tail -12 $OUT/plan_profile.txt
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_fp16.log 2>&1; tail -2 $OUT/bench_fp16.log | cut -c1-400
CPN_HALO=0 timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_fp16_nohalo.log 2>&1; tail -1 $OUT/bench_fp16_nohalo.log | cut -c1-300
CPN_PROFILE_RANGE=step timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
timeout -s KILL 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -o $OUT/prof_heads_conv python tools/run_heads_op.py heads.block.0 > $OUT/ncu_full.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:f2c_fast -s 4 -c 2 -o $OUT/prof_f2c python tools/bench_decode.py > $OUT/ncu_f2c.log 2>&1
find $OUT -size +40M -delete
ls -la $OUT | head -40
