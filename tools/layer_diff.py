"""Layer-by-layer comparison of the fp16 tensor-core plan against the strict fp32 plan on the same weights and input:
runs both op lists in lock-step (cpn_plan_run_op) and reports, after every op, max|a-b| / max|b| of the op's output.
Shows where the fp16 engine's deviation builds up.  Usage: python tools/layer_diff.py [arch] [H]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celldetection_b200 as cd  # noqa: E402
from celldetection_b200 import _lib as L  # noqa: E402
from celldetection_b200.utils.synth import synth_state_dict  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else 'CpnResNeXt101UNet'
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
x = torch.rand(1, 3, H, H, device='cuda')
models = {}
for prec in ('fp32', 'fp16'):
    m = getattr(cd.models, arch)(3, precision=prec)
    m.load_state_dict(synth_state_dict(m._spec, seed=0))
    models[prec] = m.cuda()
ps, pf = models['fp32']._plan(1, H, H), models['fp16']._plan(1, H, H)
outs_s, outs_f = ps.new_outputs(), pf.new_outputs()
rows = []
for i, (os_, of) in enumerate(zip(ps.g.ops, pf.g.ops)):
    ps.run_op(i, x, L.IN_F32_NCHW, outs_s)
    if i not in pf.fused:
        pf.run_op(i, x, L.IN_F32_NCHW, outs_f)
    torch.cuda.synchronize()
    if os_.dst.f32 or os_.kind == 'prep' or (i + 1 < len(pf.g.ops) and (i + 1) in pf.fused and os_.kind == 'conv'):
        continue   # bound outputs compared at the end; prep layouts differ (im2col); fused conv writes no mid tensor
    a, b = pf.read_tensor(of.dst), ps.read_tensor(os_.dst)
    err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
    rms = float(((a - b).pow(2).mean().sqrt()) / b.pow(2).mean().sqrt().clamp_min(1e-12))
    rows.append(dict(i=i, kind=os_.kind, name=os_.name, c=os_.dst.c, h=os_.dst.h, max_rel=err, rms_rel=rms,
                     absmax=float(b.abs().max())))
for name, a, b in zip(('scores', 'locfou', 'refinement'), outs_f, outs_s):
    rows.append(dict(i=-1, kind='output', name=name, c=a.shape[-1], h=a.shape[1],
                     max_rel=float((a - b).abs().max() / b.abs().max()), rms_rel=float(((a - b).pow(2).mean().sqrt()) / b.pow(2).mean().sqrt()),
                     absmax=float(b.abs().max())))
print(f'# {arch} 1x3x{H}x{H}: fp16 engine vs fp32 engine, per op output (max|a-b|/max|b|, rms(a-b)/rms(b), max|b|)')
for r in rows:
    print(f"{r['i']:4d} {r['kind']:8s} {r['max_rel']:.2e} {r['rms_rel']:.2e} {r['absmax']:10.3f}  c{r['c']} @{r['h']}  {r['name']}")
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', f'layer_diff_{arch}_{H}.json'), 'w'))
