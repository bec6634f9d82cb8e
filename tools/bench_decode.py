"""BASELINE config C5: fouriers2contours micro-benchmark (1e6 proposals x order 16 x 128 samples), achieved HBM GB/s
against the measured copy bandwidth.  Algorithmic bytes: 16*order + 8 in, 8*samples out = 1288 B / proposal."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celldetection_b200 as cd  # noqa: E402
from celldetection_b200 import _lib as L  # noqa: E402

P, ORDER, S = 1_000_000, 16, 128
peak = 6650.
if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')):
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
        peak = json.load(f)['hbm_gbs']
lib = L.load()
g = torch.Generator(device='cuda').manual_seed(0)
# 4 rotating input/output sets (4 x 1.29 GB) so nothing is re-used out of the 126 MB L2
sets = [(torch.randn(P, ORDER, 4, device='cuda', generator=g), torch.rand(P, 2, device='cuda', generator=g) * 512,
         torch.empty(P, S, 2, device='cuda')) for _ in range(4)]
trig = cd.ops.cpn.trig_table(ORDER, S, 'cuda')


def run(i):
    f, l, o = sets[i % 4]
    L.check(lib.cpn_fouriers2contours(L.ptr(f), L.ptr(l), P, ORDER, S, L.ptr(trig), None, L.ptr(o), L.stream_ptr()))


for i in range(4):
    run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for i in range(reps):
    run(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
bytes_ = P * (16 * ORDER + 8 + 8 * S)
gbs = bytes_ / (ms / 1e3) / 1e9
print(json.dumps(dict(metric='fouriers2contours decode', proposals=P, order=ORDER, samples=S, ms=ms,
                      algorithmic_bytes=bytes_, achieved_gbs=gbs, peak_gbs=peak, frac=gbs / peak,
                      mproposals_per_s=P / ms / 1e3)))
