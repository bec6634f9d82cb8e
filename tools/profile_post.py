"""Host + device time of each stage of the post-head chain (CPN.post_flat) on the bench workload."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import celldetection_b200 as cd  # noqa: E402
from celldetection_b200 import _lib as L  # noqa: E402
from celldetection_b200.ops import cpn as O  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device('cuda')
model = cd.models.CpnResNeXt101UNet(3)
g = torch.Generator().manual_seed(0)
calib = torch.rand(1, 3, 512, 512, generator=g)


def core_fn(x, sd_):
    model.load_state_dict(sd_)
    model.to(dev)
    return {k: v.float().cpu() for k, v in model.core_forward(x.to(dev)).items()}


sd = bench.build_state_dict(core_fn, calib)
model.load_state_dict(sd)
model.to(dev)
x = torch.rand(N, 3, 512, 512, generator=g).to(dev)
plan, outs, hw = model._run_plan(x, dense=True)      # dense heads: the stand-alone decode entry point reads locfou per pixel
sc, lf, rf = outs[:3]
torch.cuda.synchronize()
lib = L.load()


def stage(name, fn, reps=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    host = (time.perf_counter() - t0) / reps * 1e3
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    print(f'{name:28s} host {host:7.3f} ms   device {a.elapsed_time(b) / reps:7.3f} ms   wall {wall:7.3f} ms')
    return out


n, h, w = sc.shape
pixels = n * h * w
ws = torch.empty((int(lib.cpn_select_workspace_bytes(pixels)),), dtype=torch.uint8, device=dev)
meta = torch.zeros((2,), dtype=torch.int64, device=dev)
st = L.stream_ptr()
thr = float(model.score_thresh)
stage('select_count', lambda: lib.cpn_select_count(L.ptr(sc), None, None, pixels, thr, L.ptr(ws), L.ptr(meta), st))
P = int(meta.tolist()[0])
stage('meta.tolist (sync)', lambda: meta.tolist())
idx = torch.empty((P,), dtype=torch.int32, device=dev)
ssc = torch.empty((P,), dtype=torch.float32, device=dev)
seg = torch.zeros((n + 1,), dtype=torch.int32, device=dev)
stage('select_write', lambda: lib.cpn_select_write(L.ptr(sc), None, None, n, h * w, thr, L.ptr(ws), L.ptr(idx), L.ptr(ssc), P, L.ptr(seg), st))
S, order = 32, 5
con = torch.empty((P, S, 2), device=dev); pro = torch.empty((P, S, 2), device=dev); box = torch.empty((P, 4), device=dev)
loc = torch.empty((P, 2), device=dev); fou = torch.empty((P, order, 4), device=dev)
trig = O.trig_table(order, S, dev)
stage('decode_refine', lambda: lib.cpn_decode_refine(L.ptr(idx), P, L.ptr(lf), 5, order, n, h, w, 512, 512, L.ptr(trig), S, L.ptr(rf), 4, None, L.ptr(con), L.ptr(pro), L.ptr(box), L.ptr(loc), L.ptr(fou), st))
stage('8 x torch.empty', lambda: [torch.empty((P, S, 2), device=dev) for _ in range(8)])
keep, counts = stage('nms_segments (py wrapper)', lambda: O.nms_segments(box, ssc, seg, n, 0.2, 50000))
stage('seg/counts tolist', lambda: (seg.tolist(), counts.tolist()))
seg_h, counts_h = seg.tolist(), counts.tolist()
sel = stage('torch.cat(keep slices)', lambda: torch.cat([keep[seg_h[i]:seg_h[i] + counts_h[i]] for i in range(n)]))
K = int(sel.numel())
dst = torch.empty((K, S, 2), device=dev)
stage('gather_rows x7', lambda: [lib.cpn_gather_rows(L.ptr(con), S * 8, L.ptr(sel), K, L.ptr(dst), st) for _ in range(7)])
stage('post_flat (all)', lambda: model.post_flat(sc, lf, rf, hw))
stage('plan.forward', lambda: plan.forward(x, L.IN_F32_NCHW), reps=5)
print('P', P, 'K', K)
