"""Tiled CPN inference over large images: the B200-native counterpart of ``celldetection_scripts.cpn_inference``
(/root/reference/celldetection_scripts/cpn_inference.py: ``apply_model`` :311-429, ``cpn_inference`` :432-869,
``TileLoader`` :51-130, ``oom_safe_gather_dict`` :257-308) and ``cd.get_tiling_slices``
(/root/reference/celldetection/util/util.py:1305-1354).

Semantics kept: tiles from ``get_tiling_slices`` (last tile shifted inward, never padded); per-tile model call with
``offsets=[w0, h0]``; removal of contours touching a tile border that is not an image border (tile-local test,
``border_removal`` px); concatenation of all tiles; one global NMS with the model's ``nms_thresh``
(``stitching_rule='nms'``).  Multi-GPU: tiles are sharded ``i -> rank i % world``; each rank keeps its detections on
its GPU; ONE padded ``all_gather`` (NCCL over NVLink) exchanges the packed detection records, every rank then runs the
same global NMS, so every rank returns the identical, complete result (the reference gathers to rank 0 with
sequential send/recv).  No Lightning, no DataLoader workers: crops are staged through pinned host memory and copied
asynchronously while the previous batch computes.
"""
import os
from collections import OrderedDict
from itertools import product
from typing import Sequence, Union

import numpy as np
import torch

from . import _lib as L
from .ops import cpn as O
from .ops.boxes import filter_by_box_voting

__all__ = ['get_tiling_slices', 'apply_model', 'cpn_inference', 'shard_items']


def get_tiling_slices(size: Sequence[int], crop_size: Union[int, Sequence[int]], strides: Union[int, Sequence[int]],
                      return_overlaps=False):
    """util/util.py:1305-1354 -> (iterator of slice tuples, [overlaps], tiles per dimension)."""
    assert isinstance(size, (tuple, list))
    nd = len(size)
    crop_size = (crop_size,) * nd if isinstance(crop_size, int) else tuple(crop_size)
    strides = (strides,) * nd if isinstance(strides, int) else tuple(strides)
    slices, shape, overlaps = [], [], []
    for axis in range(nd):
        if crop_size[axis] >= size[axis]:
            tl = [size[axis]]
        else:
            tl = range(crop_size[axis],
                       1 + crop_size[axis] + (int(np.ceil((size[axis] - crop_size[axis]) / strides[axis]))) *
                       strides[axis], strides[axis])
        stops = np.minimum(tl, size[axis])
        starts = np.maximum(0, stops - crop_size[axis])
        overlaps_start = np.concatenate((starts[:1], stops[:-1])) - starts
        axis_slices, axis_overlaps = [], []
        for a, b, *ov in zip(starts, stops, overlaps_start, np.concatenate((overlaps_start[1:], [0]))):
            axis_slices.append(slice(int(a), int(b)))
            axis_overlaps.append(ov)
        slices.append(axis_slices), shape.append(len(starts)), overlaps.append(axis_overlaps)
    if return_overlaps:
        return product(*slices), product(*overlaps), shape
    return product(*slices), shape


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def pack_records(res: OrderedDict):
    """Flat per-detection float32 records [K, L] of every key of ``res`` (classes are stored as float; exact for
    small ints).  Returns (records, [(key, width), ...])."""
    K = int(res['scores'].shape[0])
    # explicit widths: reshape(K, -1) cannot infer a width for K == 0 (a rank without detections)
    cols = [(k, v.reshape(K, int(np.prod(v.shape[1:], dtype=np.int64))).float()) for k, v in res.items()]
    return torch.cat([c for _, c in cols], 1).contiguous(), [(k, c.shape[1]) for k, c in cols]


def unpack_records(rec, widths, like: OrderedDict):
    out, o = OrderedDict(), 0
    for k, wd in widths:
        v = rec[:, o:o + wd]
        o += wd
        shape = (rec.shape[0],) + tuple(like[k].shape[1:])
        out[k] = v.reshape(shape).long() if k in ('classes',) else v.reshape(shape).contiguous()
    return out


def canonical_order(res: OrderedDict):
    """Sort detections by their (tile index, row within tile) key so that the input order of the global NMS -- which
    breaks score ties -- is the single-process tile order no matter how tiles were sharded over ranks."""
    key = res['order_key']
    k64 = key[:, 0].to(torch.int64) * (1 << 24) + key[:, 1].to(torch.int64)
    order = torch.argsort(k64, stable=True)
    return OrderedDict((k, v[order]) for k, v in res.items())


def allgather_detections(res: OrderedDict, group=None):
    """One collective for all keys: all_gather of the per-rank counts (one small tensor, ONE host read-back for all
    ranks) + ONE all_gather of the packed, max-count-padded record buffer; returns the concatenation in rank order on
    every rank.  Ranks without detections take part with a zero count."""
    dist, rank, world = _dist()
    if dist is None or world == 1:
        return res
    rec, widths = pack_records(res)
    dev = rec.device
    cnts = torch.zeros((world,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnts, torch.tensor([rec.shape[0]], dtype=torch.int64, device=dev), group=group)
    counts = cnts.tolist()                               # the exchange's only host synchronisation
    mx = max(max(counts), 1)
    padded = torch.zeros((mx, rec.shape[1]), dtype=torch.float32, device=dev)
    padded[:rec.shape[0]] = rec
    gathered = torch.empty((world * mx, rec.shape[1]), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(gathered, padded, group=group)   # the one payload collective (NCCL over NVLink)
    parts = [gathered[r * mx:r * mx + counts[r]] for r in range(world)]
    return unpack_records(torch.cat(parts, 0), widths, res)


def _to_rgb(img):
    if img.ndim == 2:
        img = img[..., None]
    if img.shape[-1] == 1:
        img = np.repeat(img, 3, axis=-1)   # cv2.COLOR_GRAY2RGB (cpn_inference.py:330-331)
    return img


def _tile_bounds(mask, point_mask, point_mask_exclusive, sl):
    """TileLoader.__getitem__ (cpn_inference.py:93-111): score bounds of one tile, or ``False`` if the tile is skipped
    (empty mask / point-mask crop).  Returns (upper [th,tw,1] | None, lower | None)."""
    upper = lower = None
    if mask is not None:
        crop = mask[sl]
        if not np.any(crop):
            return False
        upper = (crop[..., None] if crop.ndim == 2 else crop).astype('float32')
    if point_mask is not None:
        crop = point_mask[sl]
        if not np.any(crop):
            return False
        lower = np.clip(crop[..., None] if crop.ndim == 2 else crop, 0., 1.).astype('float32')
        if point_mask_exclusive:
            upper = lower
    return upper, lower


def _apply_single(img, model, mask, point_mask, point_mask_exclusive, crop_size, strides, batch_size, border_removal,
                  rules, stitching_rule, nms_thresh, dev, timings=None, reps=1, transforms=None):
    """One model over all tiles of ``img`` (this rank's share), border removal, exchange, stitch NMS
    (cpn_inference.py:354-411).  ``img``: ``Array[h, w, c]`` on the host (crops are staged through double-buffered pinned
    memory by a helper thread) or a uint8 / float32 CUDA tensor ``[h, w, c]`` already resident on ``dev`` (crops are device
    slices).  The per-tile border test only marks rows; compaction, the order keys and the exchange run ONCE after the
    last batch, so the tile loop has no host synchronisation besides the model's own.

    Test-time repetitions (TileLoader, cpn_inference.py:85-91,114-118): every tile is one work item per ``rep_idx`` in
    ``range(reps)`` (item = tile * reps + rep, dealt round-robin to the ranks); ``transforms(crop, rep_idx) -> (crop, meta)``
    changes the INPUT crop on the host only -- the reference never maps detections back (``meta`` just travels with the batch),
    all repetitions' detections meet in the global NMS."""
    on_device = isinstance(img, torch.Tensor)
    H, W = int(img.shape[0]), int(img.shape[1])
    slices, overlaps, (h_tiles, w_tiles) = get_tiling_slices((H, W), tuple(crop_size), tuple(strides),
                                                             return_overlaps=True)
    slices, overlaps = list(slices), list(overlaps)
    ex_br = float(stitching_rule != 'nms' and 'ex_br' in rules)
    dist, rank, world = _dist()
    bounds = {}
    todo = list(range(len(slices)))
    if mask is not None or point_mask is not None:      # tiles with an empty (point-)mask crop are never inferred
        todo = []
        for t in range(len(slices)):
            bnd = _tile_bounds(mask, point_mask, point_mask_exclusive, slices[t])
            if bnd is not False:
                bounds[t] = bnd
                todo.append(t)
    mine = shard_items(todo, reps, rank, world)
    th, tw = (slices[0][0].stop - slices[0][0].start), (slices[0][1].stop - slices[0][1].start)
    is_u8 = img.dtype in (np.uint8, torch.uint8)
    C = int(img.shape[-1])
    lib = L.load()
    main = torch.cuda.current_stream(dev)
    nb = (len(mine) + batch_size - 1) // batch_size
    acc = []          # per batch: (flat dict, keep mask [K] uint8, tile id of each row [K] float, K)
    pool = None
    t_start = _now(dev, timings)

    if not on_device:
        # double-buffered pinned staging; crops are staged by a helper thread (numpy slicing + the pinned copy release
        # the GIL) while the main thread drives the GPU: model calls block on the proposal count, so same-thread staging
        # would serialise with the GPU
        stage = [torch.empty((batch_size, th, tw, C), dtype=torch.uint8 if is_u8 else torch.float32).pin_memory()
                 for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        cur_dev = torch.cuda.current_device()

        def load(bi, slot):
            torch.cuda.set_device(cur_dev)
            ids = mine[bi * batch_size:(bi + 1) * batch_size]
            buf = stage[slot]
            for j, it in enumerate(ids):
                crop = img[slices[it // reps]]
                if transforms is not None:
                    crop = _checked_transform(transforms, crop, it % reps)
                buf[j].copy_(torch.from_numpy(np.ascontiguousarray(crop)))
            with torch.cuda.stream(copy_stream):
                d = buf[:len(ids)].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return ids, d, ev

        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=1)
    tiles_range = L.nvtx_range('cpn.tiles').__enter__()
    try:
        nxt = pool.submit(load, 0, 0) if (pool is not None and nb) else None
        for bi in range(nb):
            if on_device:
                ids = mine[bi * batch_size:(bi + 1) * batch_size]
                d = torch.stack([img[slices[it // reps]] for it in ids], 0)
            else:
                ids, d, ev = nxt.result()
                main.wait_event(ev)
                d.record_stream(main)        # allocated on the copy stream, consumed on the main stream
                if bi + 1 < nb:
                    # the other staging slot was consumed two batches ago (its copy finished before the previous forward)
                    nxt = pool.submit(load, bi + 1, (bi + 1) % 2)
            tiles = [it // reps for it in ids]
            offs = torch.tensor([[slices[t][1].start, slices[t][0].start] for t in tiles], dtype=torch.float32).to(
                dev, non_blocking=True)
            kw = dict(offsets=offs)
            if bounds:       # [B,1,th,tw] score bounds at input resolution (resized by the model, cpn.py:118-123)
                for j, name in ((0, 'scores_upper_bound'), (1, 'scores_lower_bound')):
                    if bounds[tiles[0]][j] is not None:
                        b_np = np.stack([bounds[t][j][..., 0] for t in tiles], 0)[:, None]
                        kw[name] = torch.from_numpy(np.ascontiguousarray(b_np)).to(dev)
            if is_u8:
                flat, counts = model.forward_flat(d, L.IN_U8_NHWC, **kw)
            else:
                flat, counts = model.forward_flat(d.permute(0, 3, 1, 2).contiguous(), L.IN_F32_NCHW, **kw)
            K = int(sum(counts))
            if K == 0:
                continue
            meta = []
            for t in tiles:
                h_i, w_i = np.unravel_index(t, (h_tiles, w_tiles))
                (_, ov_y1), (_, ov_x1) = overlaps[t]      # right/bottom overlaps (cpn_inference.py:382-385)
                meta.append([slices[t][1].start, slices[t][0].start, th, tw, float(h_i > 0), float(w_i < w_tiles - 1),
                             float(h_i < h_tiles - 1), float(w_i > 0), ex_br, float(tw - ov_x1), float(th - ov_y1), 0.])
            meta = torch.tensor(meta, dtype=torch.float32).to(dev, non_blocking=True)
            tile_of_row = torch.repeat_interleave(torch.arange(len(ids), dtype=torch.int32),
                                                  torch.tensor(counts)).to(dev, non_blocking=True)
            keep = torch.empty((K,), dtype=torch.uint8, device=dev)
            L.check(lib.cpn_border_filter(L.ptr(flat['contours']), L.ptr(tile_of_row), L.ptr(meta), K,
                                          int(flat['contours'].shape[1]), float(border_removal), L.ptr(keep),
                                          L.stream_ptr()), 'border_filter')
            tile_ids = torch.repeat_interleave(torch.tensor(ids, dtype=torch.float32), torch.tensor(counts))
            rows = torch.arange(K, dtype=torch.float32)
            acc.append((flat, keep, torch.stack((tile_ids, rows), 1).to(dev, non_blocking=True)))
    finally:
        tiles_range.__exit__()
        if pool is not None:
            pool.shutdown(wait=True)
    t_tiles = _now(dev, timings)
    if acc:
        keep_all = torch.cat([a[1] for a in acc], 0)
        sel = torch.nonzero(keep_all, as_tuple=False).reshape(-1)          # the one compaction of this rank's detections
        res = OrderedDict((k, torch.cat([a[0][k] for a in acc], 0)[sel]) for k in acc[0][0].keys())
        res['order_key'] = torch.cat([a[2] for a in acc], 0)[sel]          # (global tile index, row in batch)
    else:
        S, order = int(model.samples), int(min(model.order, model.core_order))
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)  # noqa: E731
        res = OrderedDict(contours=z(0, S, 2), boxes=z(0, 4), scores=z(0), classes=z(0, dt=torch.long),
                          locations=z(0, 2), fourier=z(0, order, 4), contour_proposals=z(0, S, 2))
        if getattr(model, 'uncertainty_head', False):
            res['box_uncertainties'] = z(0, 4)
        res['order_key'] = z(0, 2)
    with L.nvtx_range('cpn.exchange'):
        res = allgather_detections(res)
        if world > 1:
            res = canonical_order(res)      # 1-GPU and N-GPU runs feed the global NMS the same sequence
        res.pop('order_key')
    t_gather = _now(dev, timings)
    if 'nms' in rules and res['boxes'].shape[0] > 0:
        with L.nvtx_range('cpn.stitch'):
            keep = O.nms(res['boxes'], res['scores'], nms_thresh)
            res = OrderedDict((k, v[keep]) for k, v in res.items())
    if timings is not None:
        t_end = _now(dev, timings)
        timings.update(tiles_s=t_tiles - t_start, exchange_s=t_gather - t_tiles, stitch_s=t_end - t_gather,
                       tiles_mine=len(mine), tiles_total=len(slices))
    return res


def shard_items(tiles, reps, rank, world):
    """Work items of one rank: item = tile * reps + rep_idx (TileLoader.__getitem__, cpn_inference.py:88-91: slice_idx =
    item // reps, rep_idx = item % reps) over the tiles that are inferred at all, dealt round-robin.  The item index is also
    the canonical order key of the exchange, so the union over ranks, sorted, is the single-process sequence."""
    items = [t * reps + r for t in tiles for r in range(reps)]
    return items[rank::world]


def _checked_transform(transforms, crop, rep_idx):
    """``transforms(crop, rep_idx) -> (crop, meta)`` (cpn_inference.py:118); the batch is one dense tensor, so the crop must
    keep its shape and type."""
    out, _meta = transforms(crop, rep_idx)
    out = np.asarray(out)
    if out.shape != crop.shape or out.dtype != crop.dtype:
        raise ValueError(f'transforms must keep the crop shape and dtype: {crop.shape} {crop.dtype} -> {out.shape} {out.dtype}')
    return out


def _now(dev, timings):
    """Device-synchronised wall clock for the optional stage breakdown (no synchronisation unless timings are asked)."""
    if timings is None:
        return 0.
    import time
    torch.cuda.synchronize(dev)
    return time.perf_counter()


@torch.no_grad()
def apply_model(img, models, trainer=None, mask=None, point_mask=None, crop_size=(768, 768), strides=(384, 384),
                reps=1, transforms=None, model_kwargs_list=None, batch_size=1, num_workers=0, pin_memory=False,
                border_removal=4, min_vote=1, stitching_rule='nms', gamma=1., contrast=1., brightness=0., percentile=None,
                model_parameters=None, point_mask_exclusive=False, verbose=False, grayscale=False, device=None,
                timings=None, **kwargs):
    """cpn_inference.py:311-429.  ``img``: uint8 or float ``Array[h, w, (c)]``; ``models``: a ``CPN`` instance or a
    list of them (an ensemble: every model is run over all tiles, the concatenated detections are filtered by
    ``filter_by_box_voting(boxes, nms_thresh, min_vote)`` when ``min_vote > 1`` and de-duplicated by one more NMS,
    :417-427).  ``mask`` / ``point_mask``: ``Array[h, w]`` upper / lower score bounds; tiles whose crop is empty are
    skipped (TileLoader, :93-111).  Returns the flat dict of concatenated tensors (contours, boxes, scores, classes,
    locations, fourier, contour_proposals[, box_uncertainties][, votes]) after border removal and global NMS --
    identical on every rank when distributed.  ``img`` may also be a ``[h, w, c]`` uint8 / float32 CUDA tensor (the slide
    already resident in HBM).  ``gamma`` / ``contrast`` / ``brightness`` / ``percentile`` / ``grayscale`` and 16-bit images go
    through ``preprocess`` (:196-222) on the GPU first; ``reps`` / ``transforms``: test-time repetitions (see
    ``_apply_single``).  ``timings``: optional dict that receives a device-synchronised stage breakdown (tile loop, exchange,
    stitch) of the last model."""
    if not isinstance(models, (list, tuple)):
        models = [models]
    assert len(models) >= 1, 'Please specify at least one model.'
    assert min_vote >= 1, f'Min vote smaller than minimum: {min_vote}'
    assert len(models) >= min_vote, f'Min vote greater than number of models: {min_vote}'
    reps = int(reps)
    assert reps >= 1, f'reps smaller than minimum: {reps}'
    if transforms is not None and (mask is not None or point_mask is not None):
        raise NotImplementedError('Use of masks and transforms not supported yet.')          # cpn_inference.py:116-117
    rules = stitching_rule.split(',')
    if any(r not in ('nms', 'ex_br') for r in rules):
        raise ValueError(f'Unknown stitching rule: {stitching_rule}')
    dev = torch.device(device) if device is not None else models[0].device
    if dev.type != 'cuda':
        raise RuntimeError('apply_model needs the model on a CUDA device')
    if not isinstance(crop_size, (tuple, list)):
        crop_size = (crop_size,) * 2
    if not isinstance(strides, (tuple, list)):
        strides = (strides,) * 2
    if model_parameters:                                     # resolve_model, cpn_inference.py:246-252
        for model in models:
            for k, v in _parse_model_parameters(model_parameters):
                if not hasattr(model, k):
                    raise ValueError(f'Could not find attribute {k} in model! Please check your configuration.')
                setattr(model, k, v)
    resident = isinstance(img, torch.Tensor) and img.is_cuda
    if not isinstance(img, torch.Tensor):
        img = np.asarray(img)
    # 16-bit images: numpy uint16, or a tensor of torch.uint16 / torch.int16 holding the uint16 bit pattern (signed 16-bit
    # numpy images are rejected below like every other unsupported type)
    wide = img.dtype in ((torch.uint16, torch.int16) if isinstance(img, torch.Tensor) else (np.uint16,))
    if wide or percentile is not None or gamma != 1. or contrast != 1. or grayscale:
        # the reference conditions the whole image on the CPU before tiling (:328-329); here it goes to the GPU once and
        # continues as a device-resident uint8 slide
        from .preprocessing import preprocess
        if isinstance(img, np.ndarray) and img.dtype == np.uint16:
            img = img.view(np.int16)
        with L.nvtx_range('cpn.preprocess'):
            img = preprocess(img, gamma=gamma, contrast=contrast, brightness=brightness, percentile=percentile,
                             grayscale=grayscale, device=dev)
        resident = True
    if resident:                                             # slide already resident in HBM: [h, w, c] uint8 / float32
        if img.dim() != 3 or img.dtype not in (torch.uint8, torch.float32) or img.device != dev:
            raise ValueError('a device-resident image must be a [h, w, c] uint8 or float32 tensor on the model device')
        if img.shape[-1] == 1:
            img = img.expand(-1, -1, 3)
        if transforms is not None:
            img = img.cpu().numpy()                          # transforms are host callables on numpy crops
    else:
        img = _to_rgb(np.asarray(img))
        if img.dtype.kind == 'f':
            # NOTE: the reference percentile-normalises every non-uint8 image to uint8 (:200-202); floating-point images are
            # taken as already scaled to [0, 1] here (the model's own input contract) -- 16-bit integer images are normalised
            img = img.astype(np.float32)
        elif img.dtype != np.uint8:
            raise ValueError('image must be uint8, uint16 or floating point')
    mask = None if mask is None else np.asarray(mask)
    point_mask = None if point_mask is None else np.asarray(point_mask)
    results, nms_thresh = None, None
    for model in models:
        nms_thresh = kwargs.get('nms_thresh', model.nms_thresh)
        res = _apply_single(img, model, mask, point_mask, point_mask_exclusive, crop_size, strides, batch_size,
                            border_removal, rules, stitching_rule, nms_thresh, dev, timings=timings, reps=reps,
                            transforms=transforms)
        if results is None:
            results = res
        else:                      # keys shared by all models (an uncertainty head may be missing in some)
            results = OrderedDict((k, torch.cat((results[k], res[k]), 0)) for k in results if k in res)
    # Remove duplicates from multi model (cpn_inference.py:417-427)
    if len(models) > 1 and results['boxes'].shape[0] > 0:
        if min_vote > 1:
            keep, votes = filter_by_box_voting(results['boxes'], nms_thresh, min_vote, return_votes=True)
            results = OrderedDict((k, v[keep.long()]) for k, v in results.items())
            results['votes'] = votes
        keep = O.nms(results['boxes'], results['scores'], nms_thresh)   # nms_thresh inherited from the last model
        results = OrderedDict((k, v[keep]) for k, v in results.items())
    return results


def _parse_model_parameters(model_parameters):
    """``{'nms_thresh': 0.3}`` or the CLI form ``'nms_thresh=0.3,certainty_thresh=None,refinement=False'``
    (cpn_inference.py:583-590): values of the string form are Python literals (bool / None / numbers), anything that
    does not parse stays a string."""
    import ast
    if isinstance(model_parameters, dict):
        return [(str(k).strip(), v) for k, v in model_parameters.items()]
    out = []
    for kv in str(model_parameters).split(','):
        if not kv.strip():
            continue
        k, sep, v = kv.partition('=')
        if not sep:
            raise ValueError(f'model parameter {kv!r} is not of the form key=value')
        try:
            val = ast.literal_eval(v.strip())
        except (ValueError, SyntaxError):
            val = v.strip()
        out.append((k.strip(), val))
    return out


def _load_image_file(path, dataset='image'):
    """File inputs (cpn_inference.py:689-716): ``.h5`` / ``.hdf5`` -> the named dataset, anything else through OpenCV (the
    reference's imageio / pytiff readers are third-party packages outside this image), channels as RGB."""
    ext = os.path.splitext(path)[1].lower()
    if ext in ('.h5', '.hdf5'):
        from .utils.outputs import from_h5
        return from_h5(path, dataset)
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(f'could not read image {path}')
    if img.ndim == 3:
        img = img[..., [2, 1, 0] + ([3] if img.shape[-1] == 4 else [])]
    return np.ascontiguousarray(img)


def cpn_inference(inputs, models, outputs=None, tile_size=1024, stride=768, border_removal=4, stitching_rule='nms',
                  batch_size=1, devices='auto', precision=None, return_results=True, model_parameters=None,
                  labels=False, flat_labels=False, properties=None, spacing=1., separator='-', overlay=False,
                  inputs_dataset='image', masks=None, point_masks=None, masks_dataset='mask',
                  point_masks_dataset='point_mask', skip_existing=False, verbose=False, **kwargs):
    """Entry point in the spirit of cpn_inference.py:432-869: run tiled inference for each input (numpy arrays or image /
    hdf5 file names).  ``models``: CPN instance(s) or a filename loadable by ``load_model``; ``masks`` / ``point_masks`` /
    ``min_vote`` / ``gamma`` / ``percentile`` / ``reps`` ... are forwarded to ``apply_model``.  ``labels`` / ``flat_labels`` add
    the rasterised label image (``[h, w, c]``, contours2labels) and its channel-free form (``[h, w]``,
    resolve_label_channels) to each result (:805-818).  With ``outputs`` (a directory) every input's results are written like
    the reference's (:797-851): ``<name>.h5`` with all result tensors (+ the call's arguments as json attribute of
    ``contours``), ``<name>[_flat].csv`` region-property tables when ``properties`` are given, ``<name>_overlay.tif`` with
    ``overlay=True``; ``<name>`` is ``ndarray_<index>`` for array inputs.  In a distributed run rank 0 writes.
    ``masks`` / ``point_masks``: one array or file name per input (:595-622, 757-768), handed to ``apply_model`` as that
    input's ``mask`` / ``point_mask``.  Returns ``{index: result dict}``."""
    from .utils import load_model
    if not isinstance(inputs, (list, tuple)):
        inputs = [inputs]
    per_input = {}
    for key, lst, ds in (('mask', masks, masks_dataset), ('point_mask', point_masks, point_masks_dataset)):
        if lst is None:
            continue
        if not isinstance(lst, (list, tuple)):
            lst = [lst]
        assert len(lst) == len(inputs), (f'Expecting same number of inputs and {key}s, but found {len(inputs)} inputs and '
                                         f'{len(lst)} {key}s.')
        assert key not in kwargs, f'pass either {key}s (one per input) or {key}'
        per_input[key] = [(_load_image_file(m, ds) if isinstance(m, str) else m) for m in lst]
    args = dict(tile_size=tile_size, stride=stride, border_removal=border_removal, stitching_rule=stitching_rule,
                batch_size=batch_size, precision=precision, model_parameters=model_parameters, labels=labels,
                flat_labels=flat_labels, properties=properties, spacing=spacing, separator=separator, overlay=overlay,
                models=models if isinstance(models, str) else None,
                **{k: v for k, v in kwargs.items() if isinstance(v, (int, float, str, bool, type(None), list, tuple))})
    if isinstance(models, str):
        models = load_model(models)
    models = list(models) if isinstance(models, (list, tuple)) else [models]
    for i, model in enumerate(models):
        if model.device.type != 'cuda':
            models[i] = model = model.cuda()
        # Lightning precision strings of the reference (cpn_inference.py:446,512): '32-true' is the tensor-core engine
        # that meets the fp32 results to 1e-3, the 16-bit modes the single-pass fp16 engine
        if precision in ('32-true', '32', 'fp16f8'):
            model.precision = 'fp16f8'
        elif precision in ('16-mixed', 'bf16-mixed', 'fp16', '16-true'):
            model.precision = 'fp16'
        elif precision in ('fp32', 'fp16x3'):
            model.precision = precision
        elif precision is not None:
            raise ValueError(f'Unknown precision: {precision!r}')
        if model_parameters:
            for k, v in _parse_model_parameters(model_parameters):
                if not hasattr(model, k):
                    raise AttributeError(f'model has no parameter {k!r}')
                setattr(model, k, v)
    results = OrderedDict()
    if outputs is not None:
        os.makedirs(outputs, exist_ok=True)
    _, rank, _ = _dist()
    for i, img in enumerate(inputs):
        if isinstance(img, str):
            dst = os.path.join(outputs, os.path.splitext(os.path.basename(img))[0] + '{ext}') if outputs is not None else None
            if skip_existing and dst is not None and os.path.isfile(dst.format(ext='.h5')):
                continue
            img = _load_image_file(img, inputs_dataset)
        else:
            dst = os.path.join(outputs, f'ndarray_{i}' + '{ext}') if outputs is not None else None
        extra = {key: lst[i] for key, lst in per_input.items()}
        results[i] = y = apply_model(img, models, crop_size=tile_size, strides=stride, border_removal=border_removal,
                                     stitching_rule=stitching_rule, batch_size=batch_size, verbose=verbose, **extra, **kwargs)
        shape = tuple(img.shape[:2])
        if labels or flat_labels:                  # cpn_inference.py:805-818
            from .data import contours2labels, resolve_label_channels
            lab = contours2labels(y['contours'], shape)
            if labels:
                y['labels'] = lab
            if flat_labels:
                y['flat_labels'] = resolve_label_channels(lab)
        if dst is not None and rank == 0:          # "(is_dist and rank == 0) or not is_dist", :801
            from .utils.outputs import write_outputs
            write_outputs(dst, y, shape, args=args, labels=labels, flat_labels=flat_labels, properties=properties,
                          spacing=spacing, separator=separator, overlay=overlay)
    return results if return_results else None
