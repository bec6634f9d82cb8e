"""Device time of selected ops of the bench plan (CUDA events, each op repeated), for A/B runs under experiment switches:
  CPN_BN_1X1=128 python tools/profile_ops.py fp16f8 name[,name...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celldetection_b200 as cd  # noqa: E402
from celldetection_b200 import _lib as L  # noqa: E402
from celldetection_b200.utils.synth import synth_state_dict  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else 'fp16f8'
names = sys.argv[2].split(',') if len(sys.argv) > 2 else None
m = cd.models.CpnResNeXt101UNet(3, precision=prec)
m.load_state_dict(synth_state_dict(m._spec, seed=0))
m = m.cuda()
x = torch.rand(16, 3, 512, 512, device='cuda')
plan = m._plan(16, 512, 512)
outs = plan.new_outputs()
plan.forward(x, L.IN_F32_NCHW, outs)
torch.cuda.synchronize()
tag = ' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith('CPN_'))
tot = 0.
for i, op in enumerate(plan.g.ops):
    if i in plan.fused or (names and op.name not in names):
        continue
    reps = 10
    for _ in range(3):
        plan.run_op(i, x, L.IN_F32_NCHW, outs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.run_op(i, x, L.IN_F32_NCHW, outs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tot += ms
    if names:
        print(f'[{tag}] {op.name:45s} {ms * 1e3:9.1f} us', flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    plan.forward(x, L.IN_F32_NCHW, outs)
e0.record()
for _ in range(5):
    plan.forward(x, L.IN_F32_NCHW, outs)
e1.record()
torch.cuda.synchronize()
print(f'[{tag}] sum of listed ops {tot:.3f} ms; plan.forward {e0.elapsed_time(e1) / 5:.3f} ms', flush=True)
