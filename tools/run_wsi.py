"""BASELINE config C4: tiled inference over one large synthetic image (default 3x16384x16384 uint8, crop 512,
stride 384 -> 1849 tiles), tiles sharded over the ranks of one node, ONE padded all_gather of the packed detection
records over NCCL, then the global grid NMS on every rank.

  python tools/run_wsi.py [--size 16384] [--stride 384] [--batch 16]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_wsi.py ...

Prints one JSON line (rank 0): tiles, tiles/s (max over ranks, device-synchronised wall time), detections, and a
digest of the stitched result so that runs with different N can be compared for identical output."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celldetection_b200 as cd  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=16384)
    ap.add_argument('--crop', type=int, default=512)
    ap.add_argument('--stride', type=int, default=384)
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--arch', default='CpnResNeXt101UNet')
    ap.add_argument('--precision', default='fp16')
    ap.add_argument('--labels', action='store_true', help='also rasterise the stitched contours (contours2labels)')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    bench.ARCH = args.arch
    model = getattr(cd.models, args.arch)(3, precision=args.precision)
    g = torch.Generator().manual_seed(0)
    calib = torch.rand(1, 3, 512, 512, generator=g)

    def core_fn(x, sd_):
        model.load_state_dict(sd_)
        model.to(dev)
        return {k: v.float().cpu() for k, v in model.core_forward(x.to(dev)).items()}

    sd = bench.build_state_dict(core_fn, calib)
    model.load_state_dict(sd)
    model.to(dev)
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(args.size, args.size, 3), dtype=np.uint8)
    # warm-up (plan compilation, allocator) on a small crop
    cd.apply_model(img[:args.crop * 2, :args.crop * 2], [model], crop_size=args.crop, strides=args.stride,
                   batch_size=args.batch)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = cd.apply_model(img, [model], crop_size=args.crop, strides=args.stride, batch_size=args.batch)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    slices, shape = cd.get_tiling_slices((args.size, args.size), args.crop, args.stride)
    ntiles = shape[0] * shape[1]
    h = hashlib.sha256()
    parts = {}
    for k in ('boxes', 'scores', 'contours'):
        b = res[k].cpu().numpy().tobytes()
        h.update(b)
        parts[k] = hashlib.sha256(b).hexdigest()[:8]
    # order-independent digest: rows sorted by (score, box) -- separates "same set, different order" from "different values"
    rows = np.concatenate((res['scores'].cpu().numpy()[:, None], res['boxes'].cpu().numpy()), 1)
    parts['set'] = hashlib.sha256(rows[np.lexsort(rows.T[::-1])].tobytes()).hexdigest()[:8]
    extra = {}
    if args.labels and rank == 0:           # the step after the path (cpn_inference.py:809-813)
        cd.data.contours2labels(res['contours'][:64], (args.size, args.size))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        lab = cd.data.contours2labels(res['contours'], (args.size, args.size))
        torch.cuda.synchronize()
        extra = dict(labels_seconds=time.perf_counter() - t1, label_channels=int(lab.shape[2]),
                     labelled_pixels=int((lab > 0).sum()))
        del lab
    if rank == 0:
        print(json.dumps(dict(config='C4', arch=args.arch, size=args.size, crop=args.crop, stride=args.stride,
                              tiles=ntiles, n_gpus=world, seconds=dt, tiles_per_s=ntiles / dt,
                              detections=int(res['scores'].shape[0]), digest=h.hexdigest()[:16],
                              precision=args.precision, digests=parts, **extra)), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
