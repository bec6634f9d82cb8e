"""Per-op device time of a compiled plan (CUDA events around cpn_plan_run_op, each op repeated), with the achieved
TFLOP/s (convs) and GB/s (algorithmic activation + weight bytes) per op.  Usage:
  python tools/profile_plan.py [arch] [N] [H] [precision] > gpurun_out/plan_profile.txt"""
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
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
prec = sys.argv[4] if len(sys.argv) > 4 else 'fp16'
m = getattr(cd.models, arch)(3, precision=prec)
m.load_state_dict(synth_state_dict(m._spec, seed=0))
m = m.cuda()
x = torch.rand(N, 3, H, H, device='cuda')
plan = m._plan(N, H, H)
outs = plan.new_outputs()
plan.forward(x, L.IN_F32_NCHW, outs)
torch.cuda.synchronize()
es = 2 if prec == 'fp16' else 4
rows, tot = [], 0.
for i, op in enumerate(plan.g.ops):
    if i in plan.fused:
        continue   # runs inside the preceding convolution's epilogue
    reps = 3
    plan.run_op(i, x, L.IN_F32_NCHW, outs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.run_op(i, x, L.IN_F32_NCHW, outs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    d = op.dst
    flops = byt = 0
    if op.kind == 'conv':
        flops = 2 * N * d.h * d.w * d.c * (op.src.c // op.params.groups) * op.k * op.k
        byt = N * (op.src.h * op.src.w * op.src.c + d.h * d.w * d.c) * es + d.c * (op.src.c // op.params.groups) * op.k * op.k * es
        if op.res is not None:
            byt += N * op.res.h * op.res.w * op.res.c * es
    elif op.kind == 'proj':
        flops = 2 * N * d.h * d.w * d.c * op.cin
        byt = N * d.h * d.w * (op.cin * es + d.c * 4)
    elif op.src is not None:
        byt = N * (op.src.h * op.src.w * op.src.c + d.h * d.w * d.c) * es
    eng = ''
    if op.kind == 'conv':
        eng = 'tc' if plan.ops[i].engine == L.ENGINE_TCGEN05 else 'simt'
    rows.append(dict(i=i, kind=op.kind, eng=eng, name=op.name, cin=op.src.c if op.src else 0, cout=d.c, k=op.k,
                     stride=op.stride, g=op.params.groups if op.params else 1, h=d.h, w=d.w, ms=ms,
                     tflops=flops / ms / 1e9 if flops else 0., gbs=byt / ms / 1e6))
    tot += ms
print(f'# {arch} N={N} H={H} {prec}: sum of per-op times {tot:.3f} ms ({N / tot * 1e3:.1f} tiles/s if serial)')
for r in sorted(rows, key=lambda r: -r['ms'])[:45]:
    print(f"{r['i']:4d} {r['kind']:8s} {r['eng']:4s} {r['ms']:8.3f} ms {100 * r['ms'] / tot:5.1f}% {r['tflops']:8.1f} TF/s "
          f"{r['gbs']:8.1f} GB/s  k{r['k']} s{r['stride']} g{r['g']} {r['cin']}->{r['cout']} @{r['h']}x{r['w']}  {r['name']}")
agg = {}
for r in rows:
    key = (r['kind'], r['eng'], f"k{r['k']}s{r['stride']}g{'G' if r['g'] > 1 else '1'}")
    a = agg.setdefault(key, [0., 0])
    a[0] += r['ms']
    a[1] += 1
print('# by class')
for k, (ms, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{str(k):40s} {cnt:4d} ops {ms:8.3f} ms {100 * ms / tot:5.1f}%')
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
with open(os.path.join(ROOT, 'gpurun_out', f'plan_profile_{arch}_{N}x{H}_{prec}.json'), 'w') as f:
    json.dump(rows, f)

# ---- whole-step breakdown: plan forward (one C call) vs post-head chain vs full forward_flat ----
def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


import time
t_plan = timeit(lambda: plan.forward(x, L.IN_F32_NCHW, outs))
sc, lf, rf = outs[:3]
t_post = timeit(lambda: m.post_flat(sc, lf, rf, (H, H), plan=plan))   # (uncalibrated synthetic heads: few or no proposals)
t_full = timeit(lambda: m.forward_flat(x))
t0 = time.perf_counter()
for _ in range(5):
    plan.forward(x, L.IN_F32_NCHW, outs)
host_ms = (time.perf_counter() - t0) / 5 * 1e3
torch.cuda.synchronize()
print(f'# plan.forward {t_plan:.3f} ms (host-side enqueue {host_ms:.3f} ms), post_flat {t_post:.3f} ms, forward_flat {t_full:.3f} ms')
