"""Functional CPN ops with the reference's names and argument meaning, running on the C ABI's CUDA kernels.

Mirrors /root/reference/celldetection/ops/cpn.py: ``fouriers2contours`` (:44-95), ``rel_location2abs_location``
(:15-41), ``scale_contours`` / ``scale_fourier`` (:98-165), ``batched_box_nmsi`` (:189-227),
``remove_border_contours`` (:258-290), plus ``nms`` (torch.ops.torchvision.nms as called at :211-223).
Inputs must be CUDA tensors; there is no CPU path.
"""
import ctypes
from typing import Dict, List

import numpy as np
import torch
from torch import Tensor

from .. import _lib as L

__all__ = ['fouriers2contours', 'rel_location2abs_location', 'get_scale', 'scale_contours', 'scale_fourier',
           'batched_box_nmsi', 'batched_box_nms', 'remove_border_contours', 'filter_contours_by_stitching_rule', 'nms',
           'nms_grid', 'trig_table', 'bucket_table', 'refinement_bucket_weight', 'resolve_refinement_buckets',
           'NMS_BATCH_SIZE']

NMS_BATCH_SIZE = 50000  # ops/cpn.py:12

_trig_cache: Dict[tuple, Tensor] = {}


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('celldetection_b200 ops run on CUDA tensors only (no CPU fallback).')


def trig_table(order: int, samples: int, device) -> Tensor:
    """[2, order, samples] fp32: cos / sin of ``float(np.pi) * 2 * k * linspace(0, 1, samples)`` evaluated with the
    same torch CPU ops the reference uses (ops/cpn.py:67-81), so the decode kernels see bit-identical factors."""
    key = (int(order), int(samples), str(device))
    t = _trig_cache.get(key)
    if t is None:
        sampling = torch.linspace(0, 1.0, samples)
        c = float(np.pi) * 2 * (torch.arange(1, order + 1)[..., None]) * sampling[None, :]
        t = torch.stack((torch.cos(c), torch.sin(c)), 0).contiguous().to(device)
        if len(_trig_cache) > 64:
            _trig_cache.pop(next(iter(_trig_cache)))
        _trig_cache[key] = t
    return t


def refinement_bucket_weight(index, base_index):
    """ops/cpn.py:238-244"""
    dist = torch.abs(index + 0.5 - base_index)
    sel = dist > 1
    dist = 1. - dist
    dist[sel] = 0
    return dist


def resolve_refinement_buckets(samplings, num_buckets):
    """ops/cpn.py:247-255: the three neighbouring buckets (index, weight) of every sampling position."""
    base_index = samplings * num_buckets
    base_index_int = base_index.long()
    a, b, c = base_index_int - 1, base_index_int, base_index_int + 1
    return ((a % num_buckets, refinement_bucket_weight(a, base_index)),
            (b % num_buckets, refinement_bucket_weight(b, base_index)),
            (c % num_buckets, refinement_bucket_weight(c, base_index)))


_bucket_cache: Dict[tuple, tuple] = {}


def bucket_table(samples: int, buckets: int, device):
    """([samples, 3] int32 bucket indices, [samples, 3] fp32 weights) of the default sampling ``linspace(0, 1, samples)``
    -- ``resolve_refinement_buckets`` evaluated with the reference's torch CPU ops so the bucketed refinement kernel
    (``cpn_decode_refine_buckets``) multiplies by bit-identical weights (models/cpn.py:73-82)."""
    key = (int(samples), int(buckets), str(device))
    t = _bucket_cache.get(key)
    if t is None:
        res = resolve_refinement_buckets(torch.linspace(0, 1.0, samples), buckets)
        idx = torch.stack([i for i, _ in res], -1).to(torch.int32).contiguous().to(device)
        wts = torch.stack([w for _, w in res], -1).to(torch.float32).contiguous().to(device)
        if len(_bucket_cache) > 64:
            _bucket_cache.pop(next(iter(_bucket_cache)))
        t = _bucket_cache[key] = (idx, wts)
    return t


def fouriers2contours(fourier, locations, samples=64, sampling=None, cache: Dict[str, Tensor] = None,
                      cache_size: int = 16):
    """ops/cpn.py:44-95.  fourier ``Tensor[..., order, 4]``, locations ``Tensor[..., 2]`` -> (contours
    ``Tensor[..., samples, 2]``, sampling).  Lists are mapped element-wise like the reference (:60-63)."""
    if isinstance(fourier, (tuple, list)):
        if sampling is None:
            sampling = [sampling] * len(fourier)
        return [fouriers2contours(f, l, samples=samples, sampling=s) for f, l, s in zip(fourier, locations, sampling)]
    _require_cuda(fourier, locations, sampling)
    lib = L.load()
    order = fourier.shape[-2]
    lead = fourier.shape[:-2]
    f = fourier.reshape(-1, order, 4).contiguous().float()
    loc = locations.reshape(-1, 2).contiguous().float()
    P = f.shape[0]
    sampling_ = sampling
    if sampling is None:
        sampling_ = torch.linspace(0, 1.0, samples, device=fourier.device)
        trig, per = trig_table(order, samples, fourier.device), None
    else:
        samples = sampling.shape[-1]
        if sampling.dim() == 1:   # shared explicit sampling: build its table like the reference does
            c = float(np.pi) * 2 * (torch.arange(1, order + 1)[..., None]) * sampling.detach().cpu()[None, :]
            trig, per = torch.stack((torch.cos(c), torch.sin(c)), 0).contiguous().to(fourier.device), None
        else:
            trig, per = None, sampling.reshape(-1, samples).contiguous().float()
    out = torch.empty((P, samples, 2), dtype=torch.float32, device=fourier.device)
    L.check(lib.cpn_fouriers2contours(L.ptr(f), L.ptr(loc), P, order, samples, L.ptr(trig), L.ptr(per), L.ptr(out),
                                      L.stream_ptr()), 'fouriers2contours')
    return out.reshape(tuple(lead) + (samples, 2)), sampling_


def rel_location2abs_location(locations, cache: Dict[str, Tensor] = None, cache_size: int = 16):
    """ops/cpn.py:15-41 (trivial index arithmetic; the model path fuses it into the decode kernel)."""
    h, w = locations.shape[-2:]
    d = locations.device
    off = torch.stack((torch.arange(w, device=d)[None] + torch.zeros(h, device=d)[:, None],
                       torch.zeros(w, device=d)[None] + torch.arange(h, device=d)[:, None]), 0)
    return locations + off


def get_scale(actual_size, original_size, flip=True, dtype=torch.float):
    scale = torch.as_tensor(original_size, dtype=dtype) / torch.as_tensor(actual_size, dtype=dtype)
    return scale.flip(-1) if flip else scale


def scale_contours(actual_size, original_size, contours):
    """ops/cpn.py:106-127"""
    scale = get_scale(actual_size, original_size, flip=True)
    if isinstance(contours, Tensor):
        return contours * scale.to(contours.device)
    return [c * scale.to(c.device) for c in contours]


def scale_fourier(actual_size, original_size, fourier, location):
    """ops/cpn.py:140-165"""
    scale = get_scale(actual_size, original_size, flip=True)

    def one(fo, lo):
        s = scale.to(fo.device)
        fo = fo.clone()
        fo[..., [0, 1]] = fo[..., [0, 1]] * s[0]
        fo[..., [2, 3]] = fo[..., [2, 3]] * s[1]
        return fo, lo * s

    if isinstance(fourier, Tensor):
        return one(fourier, location)
    r = [one(f, l) for f, l in zip(fourier, location)]
    return [a for a, _ in r], [b for _, b in r]


def nms_segments(boxes: Tensor, scores: Tensor, seg_offsets: Tensor, n_segments: int, iou_threshold: float,
                 chunk: int = NMS_BATCH_SIZE):
    """Segmented NMS on flat tensors.  Returns (keep [P] int32 packed per segment at its offset, counts [S] int32)."""
    _require_cuda(boxes, scores, seg_offsets)
    lib = L.load()
    P = int(boxes.shape[0])
    keep = torch.empty((max(P, 1),), dtype=torch.int32, device=boxes.device)
    counts = torch.zeros((n_segments,), dtype=torch.int32, device=boxes.device)
    ws = torch.empty((int(lib.cpn_nms_workspace_bytes(P, n_segments)),), dtype=torch.uint8, device=boxes.device)
    L.check(lib.cpn_nms_segments(L.ptr(boxes), L.ptr(scores), L.ptr(seg_offsets), n_segments, P,
                                 float(iou_threshold), int(chunk) if chunk else 0, L.ptr(ws), L.ptr(keep),
                                 L.ptr(counts), L.stream_ptr()), 'nms_segments')
    return keep, counts


GRID_NMS_MIN = 20000   # above this many boxes a single segment goes to the parallel grid NMS


def nms_grid(boxes: Tensor, scores: Tensor, iou_threshold: float, return_rounds=False):
    """Exact greedy NMS of one large box set by parallel rounds over a spatial grid (global stitch NMS)."""
    _require_cuda(boxes, scores)
    lib = L.load()
    P = int(boxes.shape[0])
    keep = torch.empty((max(P, 1),), dtype=torch.int32, device=boxes.device)
    count = torch.zeros((1,), dtype=torch.int32, device=boxes.device)
    ws = torch.empty((int(lib.cpn_nms_grid_workspace_bytes(P)),), dtype=torch.uint8, device=boxes.device)
    rounds = ctypes.c_int(0)
    boxes, scores = boxes.contiguous().float(), scores.contiguous().float()   # named: temporaries must outlive the launch
    L.check(lib.cpn_nms_grid(L.ptr(boxes), L.ptr(scores), P,
                             float(iou_threshold), L.ptr(ws), L.ptr(keep), L.ptr(count), ctypes.byref(rounds),
                             L.stream_ptr()), 'nms_grid')
    out = keep[:int(count.item())].long()
    return (out, rounds.value) if return_rounds else out


def nms(boxes: Tensor, scores: Tensor, iou_threshold: float) -> Tensor:
    """Drop-in for ``torch.ops.torchvision.nms``: kept indices (int64) in descending score order."""
    _require_cuda(boxes, scores)
    P = int(boxes.shape[0])
    if P == 0:
        return torch.zeros((0,), dtype=torch.long, device=boxes.device)
    if P >= GRID_NMS_MIN:
        return nms_grid(boxes, scores, iou_threshold)
    seg = torch.tensor([0, P], dtype=torch.int32, device=boxes.device)
    keep, counts = nms_segments(boxes.contiguous().float(), scores.contiguous().float(), seg, 1, iou_threshold, 0)
    return keep[:int(counts.item())].long()


def batched_box_nmsi(boxes: List[Tensor], scores: List[Tensor], iou_threshold: float,
                     batch_size: int = None) -> List[Tensor]:
    """ops/cpn.py:189-227: keep indices per image (relative to that image's rows), with the chunk rule."""
    assert len(scores) == len(boxes), 'The number of score tensors must match the number of box tensors.'
    batch_size = NMS_BATCH_SIZE if batch_size is None else batch_size
    if len(boxes) == 0:
        return []
    _require_cuda(*boxes, *scores)
    sizes = [int(b.shape[0]) for b in boxes]
    offs = np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)
    dev = boxes[0].device
    fb = torch.cat([b.reshape(-1, 4).float() for b in boxes], 0).contiguous()
    fs = torch.cat([s.reshape(-1).float() for s in scores], 0).contiguous()
    seg = torch.as_tensor(offs, device=dev)
    keep, counts = nms_segments(fb, fs, seg, len(boxes), iou_threshold, batch_size)
    counts = counts.tolist()
    return [(keep[int(offs[i]):int(offs[i]) + counts[i]].long() - int(offs[i])) for i in range(len(boxes))]


def batched_box_nms(boxes: List[Tensor], scores: List[Tensor], *args, iou_threshold: float):
    """ops/cpn.py:168-186"""
    keeps = batched_box_nmsi(boxes, scores, iou_threshold, batch_size=1 << 30)
    out = ([b[k] for b, k in zip(boxes, keeps)], [s[k] for s, k in zip(scores, keeps)])
    return out + tuple([a[k] for a, k in zip(arg, keeps)] for arg in args)


def remove_border_contours(contours, size, padding=1, top=True, right=True, bottom=True, left=True, offsets=None):
    """ops/cpn.py:258-290 -> bool keep mask ``Tensor[num_contours]``."""
    _require_cuda(contours)
    lib = L.load()
    K, S = int(contours.shape[0]), int(contours.shape[1])
    keep = torch.ones((K,), dtype=torch.uint8, device=contours.device)
    if K == 0:
        return keep.bool()
    h, w = size[:2]
    off = [0., 0.] if offsets is None else [float(-v) for v in torch.as_tensor(offsets).reshape(-1)[:2].tolist()]
    meta = torch.tensor([[off[0], off[1], float(h), float(w), float(top), float(right), float(bottom), float(left),
                          0., 0., 0., 0.]], dtype=torch.float32, device=contours.device)
    tile = torch.zeros((K,), dtype=torch.int32, device=contours.device)
    contours = contours.contiguous().float()
    L.check(lib.cpn_border_filter(L.ptr(contours), L.ptr(tile), L.ptr(meta), K, S,
                                  float(padding), L.ptr(keep), L.stream_ptr()), 'border_filter')
    return keep.bool()


def filter_contours_by_stitching_rule(contours, tile_size, overlaps, rule='ex_br', offsets=None, indices=False):
    """ops/cpn.py:293-325: 'ex_br' keeps a contour unless every vertex lies in the right/bottom overlap region of the
    tile (``(contours >= (tile_size - overlaps[:, 1])[[1, 0]]).any(-1).all(-1)``)."""
    _require_cuda(contours)
    if 'ex_br' not in rule.split(','):
        raise ValueError(f'Unknown stitching rule: {rule}')
    lib = L.load()
    K, S = int(contours.shape[0]), int(contours.shape[1])
    keep = torch.ones((K,), dtype=torch.uint8, device=contours.device)
    if K > 0:
        ts = torch.as_tensor(tile_size).reshape(-1).tolist()
        ov = torch.as_tensor(overlaps).reshape(2, 2).tolist()
        stop_y, stop_x = ts[0] - ov[0][1], ts[1] - ov[1][1]
        off = [0., 0.] if offsets is None else [float(-v) for v in torch.as_tensor(offsets).reshape(-1)[:2].tolist()]
        meta = torch.tensor([[off[0], off[1], float(ts[0]), float(ts[1]), 0., 0., 0., 0., 1., float(stop_x),
                              float(stop_y), 0.]], dtype=torch.float32, device=contours.device)
        tile = torch.zeros((K,), dtype=torch.int32, device=contours.device)
        contours = contours.contiguous().float()
        L.check(lib.cpn_border_filter(L.ptr(contours), L.ptr(tile), L.ptr(meta), K, S, 0.,
                                      L.ptr(keep), L.stream_ptr()), 'border_filter')
    keep = keep.bool()
    return torch.where(keep)[0] if indices else keep
