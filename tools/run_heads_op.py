"""Runs only the dominant kernel of the bench workload (the merged 7x7 head convolution of CpnResNeXt101UNet,
batch 16 x 3x512x512) a few times inside a profiler range, for `ncu --profile-from-start off --set full`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celldetection_b200 as cd  # noqa: E402
from celldetection_b200 import _lib as L  # noqa: E402
from celldetection_b200.utils.synth import synth_state_dict  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'heads.block.0'
m = cd.models.CpnResNeXt101UNet(3)
m.load_state_dict(synth_state_dict(m._spec, seed=0))
m = m.cuda()
x = torch.rand(16, 3, 512, 512, device='cuda')
plan = m._plan(16, 512, 512)
outs = plan.new_outputs()
plan.forward(x, L.IN_F32_NCHW, outs)
idx = [i for i, o in enumerate(plan.g.ops) if o.name == name][0]
for _ in range(3):
    plan.run_op(idx, x, L.IN_F32_NCHW, outs)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    plan.run_op(idx, x, L.IN_F32_NCHW, outs)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('ok', name, idx)
