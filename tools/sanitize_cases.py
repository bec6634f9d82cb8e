"""Small launches of the kernels with the most intricate synchronisation, for `compute-sanitizer --tool memcheck|racecheck`
(SURVEY section 5): conv_halo_kernel / conv_tc_kernel in the 2-pass engine (shifted-window UMMA descriptors over one
TMA-loaded patch, mbarrier rings, the shared-memory staged epilogue), c2l_paint_kernel (persistent warps with acquire /
release flags), grid_round_kernel (fixed-point NMS rounds).  Usage: python tools/sanitize_cases.py [conv|labels|nms]..."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import celldetection_b200 as cd  # noqa: E402

which = sys.argv[1:] or ['conv', 'labels', 'nms']
if 'conv' in which:
    from gpu_conv_check import run
    for engine, ids in (('tcgen05f8', (1, 3, 5, 7, 11)), ('tcgen05', (1, 11)), ('tcgen05x3', (3,))):
        for i in ids:
            print(run(engine, i), flush=True)
if 'labels' in which:
    rng = np.random.RandomState(0)
    K, S, H, W = 300, 24, 160, 200
    t = np.linspace(0, 2 * np.pi, S, endpoint=False)
    c = rng.rand(K, 1, 2) * [W, H]
    r = rng.rand(K, 1, 1) * 10 + 3
    con = torch.from_numpy((c + r * np.stack((np.cos(t), np.sin(t)), -1)[None]).astype(np.float32)).cuda()
    lab = cd.data.contours2labels(con, (H, W))
    flat = cd.data.resolve_label_channels(lab)
    torch.cuda.synchronize()
    print('labels', tuple(lab.shape), int((lab > 0).sum()), int(flat.max()), flush=True)
if 'nms' in which:
    g = torch.Generator().manual_seed(0)
    n = 24000
    xy = torch.rand(n, 2, generator=g) * 2000
    wh = torch.rand(n, 2, generator=g) * 30 + 4
    boxes = torch.cat((xy, xy + wh), 1).cuda()
    scores = torch.rand(n, generator=g).cuda()
    keep = cd.ops.cpn.nms_grid(boxes, scores, 0.2)
    keep2 = cd.ops.cpn.nms(boxes[:3000].contiguous(), scores[:3000].contiguous(), 0.2)
    torch.cuda.synchronize()
    print('nms', int(keep.numel()), int(keep2.numel()), flush=True)
print('done', flush=True)
