"""Result files of ``cpn_inference`` (/root/reference/celldetection_scripts/cpn_inference.py:797-867): the hdf5 file with
every result tensor (``cd.to_h5``, util/util.py:1357-1399), the region-property tables as csv
(``cd.data.labels2property_table``, data/misc.py:320-345) and the overlay tif (:839-849).

The region statistics are computed on the GPU (``cpn_label_props``); file formats are written on the host: hdf5 through
h5py when it is importable, otherwise through the spec-level writer in ``h5min`` (this image ships neither h5py nor
libhdf5), tif through ``tiffmin`` (no tifffile here), csv through pandas.
"""
import json
import os
from collections import OrderedDict

import numpy as np
import torch

from .. import _lib as L
from . import h5min, tiffmin

__all__ = ['to_h5', 'from_h5', 'labels2property_table', 'label_overlay', 'dict_to_json_string', 'write_outputs',
           'SUPPORTED_PROPERTIES']


def asnumpy(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    if isinstance(v, dict):
        return type(v)((k, asnumpy(x)) for k, x in v.items())
    return v


def dict_to_json_string(input_dict):
    """util/util.py:2169-2177: the json-serialisable part of a dict."""
    out = {}
    for k, v in input_dict.items():
        try:
            json.dumps(v)
            out[k] = v
        except TypeError:
            pass
    return json.dumps(out)


def to_h5(filename, mode='w', chunks=None, compression=None, overwrite=False, driver=None, create_dataset_kw=None,
          attributes=None, **kwargs):
    """util/util.py:1357-1399: write ``{dataset_name: data}`` (+ ``attributes = {dataset_name: {name: value}}``)."""
    data = OrderedDict((k, asnumpy(v)) for k, v in kwargs.items())
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None:
        with h5py.File(filename, mode, **({} if driver is None else dict(driver=driver))) as h:
            for k, v in data.items():
                chunks_ = chunks[k] if isinstance(chunks, dict) else chunks
                if isinstance(chunks_, int) and v.ndim > 1:
                    chunks_ = tuple(np.minimum((256,) * v.ndim, v.shape))
                exists = k in h
                if overwrite and exists:
                    del h[k]
                if exists and not overwrite:
                    h[k][:] = v
                    ds = h[k]
                else:
                    ds = h.create_dataset(k, data=v, compression=compression, chunks=chunks_, **(create_dataset_kw or {}))
                if (attributes or {}).get(k):
                    ds.attrs.update(attributes[k])
        return filename
    if mode != 'w' or chunks is not None or compression is not None or driver is not None or create_dataset_kw:
        raise NotImplementedError('without h5py only plain contiguous datasets in a new file can be written '
                                  '(mode="w", no chunks / compression / driver)')
    return h5min.write(filename, data, attributes=attributes)


def from_h5(filename, *keys, **kwargs):
    """Datasets of a result file (util/util.py ``from_h5``); all of them as a dict when no key is given."""
    try:
        import h5py
        with h5py.File(filename, 'r') as h:
            out = {k: h[k][:] for k in (keys or h.keys())}
    except ImportError:
        r = h5min.read(filename)
        out = {k: r[k] for k in (keys or r.keys())}
    return out[keys[0]] if len(keys) == 1 else out


# ----------------------------------------------------------------------------------------------------------------------
# region properties
# ----------------------------------------------------------------------------------------------------------------------
# skimage.measure.regionprops names (and their pre-0.19 aliases) that the GPU statistics determine exactly
SUPPORTED_PROPERTIES = ('label', 'area', 'bbox', 'centroid', 'area_bbox', 'bbox_area', 'extent',
                        'equivalent_diameter_area', 'equivalent_diameter')


def label_stats(labels):
    """``[h, w(, c)]`` integer label image (numpy or CUDA tensor) -> per channel a dict of numpy arrays
    (label, area [px], bbox [n, 4], sum_rc [n, 2]) for the labels present, ascending."""
    t = labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(labels))
    if t.dim() == 2:
        t = t[..., None]
    t = t.to('cuda' if not t.is_cuda else t.device).to(torch.int32).contiguous()
    H, W, C = (int(v) for v in t.shape)
    lib = L.load()
    max_label = int(t.max().item()) if t.numel() else 0
    slots = C * (max_label + 1)
    area = torch.empty((slots,), dtype=torch.int32, device=t.device)
    bbox = torch.empty((slots, 4), dtype=torch.int32, device=t.device)
    sums = torch.empty((slots, 2), dtype=torch.int64, device=t.device)
    flags = torch.empty((1,), dtype=torch.int32, device=t.device)
    L.check(lib.cpn_label_props(L.ptr(t), H, W, C, max_label, L.ptr(area), L.ptr(bbox), L.ptr(sums), L.ptr(flags),
                                L.stream_ptr()), 'label_props')
    area = area.cpu().numpy().view(np.uint32).reshape(C, max_label + 1)
    bbox = bbox.cpu().numpy().reshape(C, max_label + 1, 4)
    sums = sums.cpu().numpy().reshape(C, max_label + 1, 2)
    out = []
    for z in range(C):
        idx = np.nonzero(area[z])[0]
        out.append(dict(label=idx.astype(np.int64), area=area[z][idx].astype(np.int64), bbox=bbox[z][idx].astype(np.int64),
                        sum_rc=sums[z][idx].astype(np.int64)))
    return out


def labels2property_table(labels, *properties, iter_channels=True, spacing=None, separator='-', df_kwargs=None):
    """data/misc.py:320-345 for the properties in ``SUPPORTED_PROPERTIES``: one row per region, channels concatenated in
    order (each channel's regions by ascending label, like ``regionprops``); multi-valued properties become
    ``name<separator>i`` columns.  ``spacing`` (scalar or per-axis pair) scales areas and centroids like skimage >= 0.20."""
    import pandas as pd
    if len(properties) == 1 and isinstance(properties[0], (list, tuple)):
        properties, = properties
    unknown = [p for p in properties if p not in SUPPORTED_PROPERTIES]
    if unknown:
        raise NotImplementedError(f'region properties {unknown} are outside the accelerated path; supported: '
                                  f'{SUPPORTED_PROPERTIES}')
    nd = labels.dim() if isinstance(labels, torch.Tensor) else np.asarray(labels).ndim
    if nd == 3 and not iter_channels:
        raise NotImplementedError('iter_channels=False (3-d regions) is outside the accelerated path')
    sp = (1., 1.) if spacing is None else ((float(spacing),) * 2 if np.isscalar(spacing) else tuple(float(s) for s in spacing))
    px_area = sp[0] * sp[1]
    tab = None
    for st in label_stats(labels):
        cols = OrderedDict()
        n_px = st['area'].astype(np.float64)
        area = n_px * px_area
        bb = st['bbox']
        area_bbox = ((bb[:, 2] - bb[:, 0]) * (bb[:, 3] - bb[:, 1])).astype(np.float64) * px_area
        for p in properties:
            if p == 'label':
                cols['label'] = st['label']
            elif p == 'area':
                cols['area'] = area
            elif p == 'bbox':
                for i in range(4):
                    cols[f'bbox{separator}{i}'] = bb[:, i]
            elif p == 'centroid':                    # mean of the (spacing-scaled) pixel coordinates, float64
                for i in range(2):
                    cols[f'centroid{separator}{i}'] = st['sum_rc'][:, i].astype(np.float64) * sp[i] / n_px \
                        if sp[i] != 1. else st['sum_rc'][:, i].astype(np.float64) / n_px
            elif p in ('area_bbox', 'bbox_area'):
                cols[p] = area_bbox
            elif p == 'extent':
                cols['extent'] = area / area_bbox
            elif p in ('equivalent_diameter_area', 'equivalent_diameter'):
                cols[p] = np.sqrt(4 * area / np.pi)
        tab_ = pd.DataFrame(cols, **(df_kwargs or {}))
        tab = pd.concat((tab, tab_))
    return tab


# ----------------------------------------------------------------------------------------------------------------------
# overlay
# ----------------------------------------------------------------------------------------------------------------------
def label_overlay(labels):
    """RGBA uint8 overlay of a ``[h, w(, c)]`` label image.  The reference colours labels with ``cd.label_cmap(...,
    ubyte=True)``, whose default palette is RANDOM (visualization/cmaps.py:43-48), so there is no value to match: every label
    gets a fixed pseudo-random opaque colour (a hash of the label), background stays transparent, later channels win."""
    lab = asnumpy(labels)
    if lab.ndim == 2:
        lab = lab[..., None]
    out = np.zeros(lab.shape[:2] + (4,), dtype=np.uint8)
    for z in range(lab.shape[2]):
        l = lab[..., z].astype(np.uint32)
        h = (l * np.uint32(2654435761)) & np.uint32(0xffffffff)
        rgb = np.stack(((h >> 8) & 255, (h >> 16) & 255, (h >> 24) & 255), -1).astype(np.uint8) | np.uint8(64)
        m = l > 0
        out[m, :3] = rgb[m]
        out[m, 3] = 255
    return out


# ----------------------------------------------------------------------------------------------------------------------
# the per-input output step of cpn_inference
# ----------------------------------------------------------------------------------------------------------------------
def write_outputs(dst, y, image_shape, args=None, labels=False, flat_labels=False, properties=None, spacing=1.,
                  separator='-', overlay=False):
    """cpn_inference.py:797-851 for one input.  ``dst``: path with an ``{ext}`` placeholder; ``y``: result dict of CUDA
    tensors (may already hold ``labels`` / ``flat_labels``).  Adds what it computed to ``y`` and returns the dict of written
    files."""
    from ..data import contours2labels, resolve_label_channels
    do_props = properties is not None and len(properties)
    do_labels = do_props or labels or flat_labels or overlay
    labels_ = y.get('labels')
    flat_ = y.get('flat_labels')
    if do_labels and labels_ is None:
        labels_ = contours2labels(y['contours'], image_shape[:2])
    if flat_labels and flat_ is None:
        flat_ = resolve_label_channels(labels_)
    output = OrderedDict((k, v) for k, v in y.items() if k not in ('labels', 'flat_labels'))
    if labels:
        y['labels'] = output['labels'] = labels_
    if flat_labels:
        y['flat_labels'] = output['flat_labels'] = flat_
    files = OrderedDict()
    files['h5'] = dst.format(ext='.h5')
    to_h5(files['h5'], **asnumpy(output), attributes=dict(contours=dict(args=dict_to_json_string(args or {}))))
    if do_props:
        if flat_labels:
            tab = labels2property_table(flat_, properties, spacing=spacing, separator=separator)
            y['properties_flat'] = tab
            files['properties_flat'] = dst.format(ext='_flat.csv')
            tab.to_csv(files['properties_flat'])
        if labels or not flat_labels:
            tab = labels2property_table(labels_, properties, spacing=spacing, separator=separator)
            y['properties'] = tab
            files['properties'] = dst.format(ext='.csv')
            tab.to_csv(files['properties'])
    if overlay:
        vis = label_overlay(labels_)
        files['overlay'] = dst.format(ext='_overlay.tif')
        tiffmin.imwrite(files['overlay'], vis, compression='ZLIB', bigtiff=vis.size > (2 ** 28))
        y['overlay'] = vis
    y['files'] = files
    return files
