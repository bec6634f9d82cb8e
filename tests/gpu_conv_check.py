"""Stand-alone per-layer conv parity runner (spawned by test_gpu_conv.py with a timeout so that a dead-locked kernel
cannot hang the suite).  Usage: python tests/gpu_conv_check.py <engine> [case indices...]; prints one JSON per case."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from celldetection_b200.ops.conv import conv2d  # noqa: E402

# (cin, cout, k, stride, groups, n, h, w, residual, relu, bias)   -- the layer types of SURVEY appendix C.4
CASES = [
    (64, 64, 1, 1, 1, 1, 16, 16, None, False, True),          # 0 1x1
    (64, 64, 3, 1, 1, 1, 16, 16, None, True, True),           # 1 3x3 s1
    (128, 256, 3, 1, 1, 2, 24, 40, None, True, True),         # 2 3x3 s1, ragged spatial (not multiples of 16x8)
    (64, 64, 7, 1, 1, 1, 32, 32, None, True, True),           # 3 7x7 head
    (256, 768, 7, 1, 1, 1, 32, 32, None, True, True),         # 4 merged heads 256 -> 3*256
    (256, 512, 1, 2, 1, 1, 32, 32, None, False, False),       # 5 1x1 s2 downsample
    (128, 128, 3, 2, 1, 2, 32, 32, None, True, False),        # 6 3x3 s2 dense (ResNet18)
    (256, 256, 3, 1, 32, 1, 32, 32, None, True, False),       # 7 grouped 8 ch/group
    (512, 512, 3, 2, 32, 1, 32, 32, None, True, False),       # 8 grouped 16 ch/group, stride 2
    (1024, 1024, 3, 1, 32, 1, 16, 16, None, True, False),     # 9 grouped 32 ch/group
    (2048, 2048, 3, 1, 32, 1, 8, 8, None, True, False),       # 10 grouped 64 ch/group
    (256, 256, 1, 1, 1, 1, 32, 32, 'same', True, False),      # 11 bottleneck conv3 + residual + relu
    (128, 256, 1, 1, 1, 1, 32, 32, 'half', False, True),      # 12 FPN lateral + nearest-upsampled residual
    (320, 256, 3, 1, 1, 1, 32, 32, None, True, True),         # 13 decoder cat input 320
    (3072, 2048, 3, 1, 1, 1, 8, 8, None, True, True),         # 14 deepest decoder conv, long K
    (64, 128, 3, 1, 1, 1, 19, 21, None, False, True),         # 15 odd sizes
    (3, 64, 7, 2, 1, 1, 64, 64, None, True, False),           # 16 stem (SIMT only)
    (3, 64, 3, 1, 1, 1, 32, 32, None, True, True),            # 17 U22 first conv (SIMT only)
]


def run(engine, idx):
    cin, cout, k, stride, groups, n, h, w, res, relu, bias = CASES[idx]
    g = torch.Generator().manual_seed(idx)
    half = engine not in ('simt32', 'tcgen05x3', 'tcgen05f8')
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin // groups, k, k, generator=g) * (2. / (cin // groups * k * k)) ** .5
    b = torch.randn(cout, generator=g) * 0.1 if bias else None
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    r = None
    if res == 'same':
        r = torch.randn(n, cout, ho, wo, generator=g)
    elif res == 'half':
        r = torch.randn(n, cout, (ho + 1) // 2, (wo + 1) // 2, generator=g)
    if half:  # the kernels see fp16-rounded operands; give the reference the same values
        x, wt = x.half().float(), wt.half().float()
        r = None if r is None else r.half().float()
    ref = F.conv2d(x, wt, b, stride=stride, padding=pad, groups=groups)
    if r is not None:
        rr = r if res == 'same' else F.interpolate(r, size=(ho, wo), mode='nearest')
        ref = ref + rr
    if relu:
        ref = F.relu(ref)
    eng = engine if engine in ('tcgen05', 'tcgen05x3', 'tcgen05f8') else 'simt'
    out = conv2d(x.cuda(), wt.cuda(), None if b is None else b.cuda(), stride=stride, padding=pad, groups=groups,
                 residual=None if r is None else r.cuda(), relu=relu, engine=eng, half=half)
    torch.cuda.synchronize()
    out = out.cpu()
    if os.environ.get('CPN_DUMP'):
        import numpy as np
        np.savez_compressed(os.path.join(os.environ['CPN_DUMP'], f'conv_dump_{engine}_{idx}.npz'), out=out.numpy(),
                            ref=ref.numpy())
    err = float((out - ref).abs().max() / ref.abs().max().clamp_min(1e-6))
    bad = (out - ref).abs() > 2e-2 * ref.abs().max()
    return dict(case=idx, engine=engine, shape=list(CASES[idx][:8]), rel_err=err, n_bad=int(bad.sum()),
                n=int(bad.numel()), nan=bool(torch.isnan(out).any()))


if __name__ == '__main__':
    engine = sys.argv[1]
    ids = [int(a) for a in sys.argv[2:]] or list(range(len(CASES)))
    for i in ids:
        try:
            print(json.dumps(run(engine, i)), flush=True)
        except Exception as e:  # noqa
            print(json.dumps(dict(case=i, engine=engine, error=str(e)[:400])), flush=True)
