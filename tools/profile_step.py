"""Where a bench step goes: plan.forward alone vs the whole forward_flat (calibrated heads, the bench workload), device time
by CUDA events and host wall clock, sparse and dense heads.  Usage: python tools/profile_step.py [precision]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from celldetection_b200 import _lib as L  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else 'fp16f8'
dev = torch.device('cuda')
model, sd = bench.make_model(bench.ARCH, prec, dev)
xs = bench.synthetic_batches(4, dev, 1)


def timed(fn, reps=10):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(reps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, (time.perf_counter() - t0) / reps * 1e3


for sparse in (True, False):
    model.sparse_heads = sparse
    plan = model._plan(bench.BATCH, bench.TILE, bench.TILE)
    outs = plan.new_outputs()
    dev_plan, wall_plan = timed(lambda i: plan.forward(xs[i % 4], L.IN_F32_NCHW, outs))
    dev_full, wall_full = timed(lambda i: model.forward_flat(xs[i % 4]))
    l0 = L.launch_count()
    flat, counts = model.forward_flat(xs[0])
    n_launch = L.launch_count() - l0
    print(f'[{prec} sparse_heads={sparse}] plan.forward {dev_plan:.3f} ms ({plan.n_launches} launches); forward_flat {dev_full:.3f} ms '
          f'device / {wall_full:.3f} ms wall ({n_launch} library launches, {sum(counts)} kept): post-head chain + host '
          f'{dev_full - dev_plan:.3f} ms', flush=True)
