#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
HALO_CASES="1 2 3 4 13 14 15"
check() { python - "$1" <<'PY'
import json,sys
rows=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')]
ok = rows and all(('error' not in r) and r['rel_err'] < 2e-3 for r in rows)
print('PASS' if ok else 'FAIL', [(r['case'], round(r.get('rel_err', -1), 5) if 'rel_err' in r else r.get('error')) for r in rows])
sys.exit(0 if ok else 1)
PY
}
MODE=""
CPN_HALO_ALL=1 CPN_HALO_SW128=1 CPN_HALO_BASEOFF=0 timeout -s KILL 300 python tests/gpu_conv_check.py tcgen05 $HALO_CASES > $OUT/halo_sw128_b0.log 2>&1
if check $OUT/halo_sw128_b0.log; then echo "sw128 baseoff=0 OK"; MODE="CPN_HALO_SW128=1"; fi
CPN_HALO_ALL=1 CPN_HALO_SW128=1 CPN_HALO_BASEOFF=1 timeout -s KILL 300 python tests/gpu_conv_check.py tcgen05 $HALO_CASES > $OUT/halo_sw128_b1.log 2>&1
if check $OUT/halo_sw128_b1.log; then echo "sw128 baseoff=1 OK"; [ -z "$MODE" ] && MODE="CPN_HALO_SW128=1 CPN_HALO_BASEOFF=1"; fi
echo "MODE=$MODE" | tee $OUT/halo_mode.txt
echo "== default (planes, 7x7 + bn<=128)"; timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile_default.txt 2>&1; head -14 $OUT/plan_profile_default.txt
if [ -n "$MODE" ]; then
  echo "== $MODE"; env $MODE timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile_sw128.txt 2>&1; head -14 $OUT/plan_profile_sw128.txt
  echo "== $MODE HALO_ALL"; env $MODE CPN_HALO_ALL=1 timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile_sw128_all.txt 2>&1; head -14 $OUT/plan_profile_sw128_all.txt
  env $MODE timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_sw128.log 2>&1; tail -1 $OUT/bench_sw128.log | cut -c1-260
  env $MODE CPN_HALO_ALL=1 timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_sw128_all.log 2>&1; tail -1 $OUT/bench_sw128_all.log | cut -c1-260
fi
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_default.log 2>&1; tail -1 $OUT/bench_default.log | cut -c1-260
find $OUT -size +40M -delete
