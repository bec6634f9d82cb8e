#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 300 python tests/gpu_conv_check.py tcgen05 0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 > $OUT/conv_checks.log 2>&1; python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/conv_checks.log') if l.startswith('{')]
print('conv cases', len(rows), 'max rel', max(r.get('rel_err', 9) for r in rows), [r for r in rows if 'error' in r])
PY
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile.txt 2>&1; head -16 $OUT/plan_profile.txt; tail -11 $OUT/plan_profile.txt
echo "== HALO_ALL"; CPN_HALO_ALL=1 timeout -s KILL 600 python tools/profile_plan.py CpnResNeXt101UNet 16 512 fp16 > $OUT/plan_profile_all.txt 2>&1; head -14 $OUT/plan_profile_all.txt
timeout -s KILL 900 python bench.py > $OUT/bench_fp16.log 2>&1; tail -1 $OUT/bench_fp16.log | cut -c1-330
find $OUT -size +40M -delete
