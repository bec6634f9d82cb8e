"""Runs single ops of the bench workload (CpnResNeXt101UNet, batch 16 x 3x512x512) inside a profiler range, for
`ncu --profile-from-start off --set full`:  run_heads_op.py name[,name...] [precision] [arch] [batch]   (default: the dominant kernel,
the merged 7x7 head convolution) or  run_heads_op.py post  (one whole post-head chain on calibrated heads)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celldetection_b200 as cd  # noqa: E402
from celldetection_b200 import _lib as L  # noqa: E402
from celldetection_b200.utils.synth import synth_state_dict  # noqa: E402

names = (sys.argv[1] if len(sys.argv) > 1 else 'heads.block.0').split(',')
prec = sys.argv[2] if len(sys.argv) > 2 else 'fp16'
arch = sys.argv[3] if len(sys.argv) > 3 else 'CpnResNeXt101UNet'
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 16
m = getattr(cd.models, arch)(3, precision=prec)
m.load_state_dict(synth_state_dict(m._spec, seed=0))
m = m.cuda()
x = torch.rand(batch, 3, 512, 512, device='cuda')
plan = m._plan(batch, 512, 512)
outs = plan.new_outputs()
plan.forward(x, L.IN_F32_NCHW, outs)
if names == ['post']:
    # the whole post-head chain (select / decode+refine / NMS / gathers) of one step, for `ncu -k regex:...`
    from celldetection_b200.utils.synth import calibrate_heads_
    sd = synth_state_dict(m._spec, seed=0)
    calibrate_heads_(sd, lambda xx, s_: (m.load_state_dict(s_), {k: v.float().cpu() for k, v in m.core_forward(xx.cuda()).items()})[1],
                     x[:2].cpu(), fg_fraction=0.02)
    m.load_state_dict(sd)
    m = m.cuda()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print('ok post, kept', sum(len(s) for s in out['scores']))
    sys.exit(0)
for name in names:
    idx = [i for i, o in enumerate(plan.g.ops) if o.name == name][0]
    for _ in range(3):
        plan.run_op(idx, x, L.IN_F32_NCHW, outs)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    plan.run_op(idx, x, L.IN_F32_NCHW, outs)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print('ok', name, idx)
