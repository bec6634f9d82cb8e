"""contours2labels micro-benchmark: 1e5 contours x 128 samples on a 16384 x 16384 label image (the C4 slide),
GPU (cpn_contours2labels, device-resident contours, CUDA events) next to the CPU port (oracle/c2l_oracle.py, the
restated cv2 fill + greedy channel rule) on a bounded sample.  Prints one JSON line.  Lives under tests/ because it executes
the oracle as its CPU baseline (test infrastructure; not collected by pytest).  Usage: python tests/bench_c2l.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import celldetection_b200 as cd  # noqa: E402
import c2l_oracle as c2l  # noqa: E402
from test_gpu_labels import synth_contours  # noqa: E402

rng = np.random.RandomState(11)
H = W = 16384
K, S = 100000, 128
base = synth_contours(rng, 500, S, 64, 64, (4., 20.)) - 32.
centres = np.stack((rng.rand(K) * W, rng.rand(K) * H), -1).astype(np.float32)
con = (base[rng.randint(0, 500, K)] + centres[:, None]).astype(np.float32)
t = torch.from_numpy(con).cuda()
lab = cd.data.contours2labels(t, (H, W))
torch.cuda.synchronize()
times = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lab = cd.data.contours2labels(t, (H, W))
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = float(np.median(times))
# CPU port on a bounded sample: the first 2000 contours of the same set, cropped canvas not needed (same image size)
n_cpu = 2000                                   # same shapes at the same density: 2000 contours on a 2318 x 2318 image
side = int(round(H * (n_cpu / K) ** 0.5))
con_cpu = (base[rng.randint(0, 500, n_cpu)] + (rng.rand(n_cpu, 1, 2) * side).astype(np.float32)).astype(np.float32)
t0 = time.time()
c2l.contours2labels(con_cpu.copy(), (side, side), clip=True)
cpu_s = time.time() - t0
try:                                           # the third-party fill the reference calls (OpenCV), same sample, for scale
    import cv2
    t0 = time.time()
    for c in np.round(con_cpu).astype(np.int32):
        a = np.zeros((c[:, 1].max() - c[:, 1].min() + 1, c[:, 0].max() - c[:, 0].min() + 1), np.int32)
        cv2.drawContours(a, [c.reshape(-1, 1, 2)], 0, 1, -1, offset=(int(-c[:, 0].min()), int(-c[:, 1].min())))
    cv2_s = time.time() - t0
except Exception:
    cv2_s = None
painted = int((lab > 0).sum())
print(json.dumps(dict(metric='contours2labels contours/s (16384x16384, 1e5 contours x 128 samples)', value=K / (ms / 1e3),
                      unit='contours/s', ms=ms, channels=int(lab.shape[2]), painted_pixels=painted,
                      label_image_gb=lab.numel() * 4 / 1e9,
                      cpu_port=dict(value=n_cpu / cpu_s, unit='contours/s', sample=f'{n_cpu} contours on {side}x{side} (same density)',
                                    cores=1, kind='port'),
                      opencv_fill_only_contours_per_s=(n_cpu / cv2_s if cv2_s else None),
                      reference_note='reference docstring: ~137 ms for 1284 contours x 128 points on 1000x1000 = 9.4e3 contours/s '
                                     '(data/cpn.py:298)')))
