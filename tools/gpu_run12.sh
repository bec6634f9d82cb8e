#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; OUT=gpurun_out
timeout -s KILL 300 python tests/gpu_conv_check.py tcgen05x3 0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 > $OUT/conv_checks_x3.log 2>&1; cut -c1-160 $OUT/conv_checks_x3.log
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -12 $OUT/pytest.log
timeout -s KILL 900 python bench.py --precision fp16 > $OUT/bench_fp16.log 2>&1; tail -1 $OUT/bench_fp16.log | cut -c1-330
python - <<'PY' > gpurun_out/bench_x3.log 2>&1
import sys, torch, time
sys.path.insert(0, '.')
import bench, celldetection_b200 as cd
from celldetection_b200.utils.synth import synth_state_dict
m = cd.models.CpnResNeXt101UNet(3, precision='fp16x3')
m.load_state_dict(synth_state_dict(m._spec, seed=0)); m = m.cuda()
x = torch.rand(16, 3, 512, 512, device='cuda')
for _ in range(3): m.forward_flat(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): m.forward_flat(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('fp16x3 C3 batch16: %.2f ms/step -> %.1f tiles/s' % (ms, 16 / ms * 1e3))
PY
cat $OUT/bench_x3.log | tail -3
find $OUT -size +40M -delete
