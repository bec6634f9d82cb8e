"""``cd.data`` functions that follow the CPN inference path, executed by the C ABI's CUDA kernels.

``contours2labels`` mirrors /root/reference/celldetection/data/cpn.py:292-358 (and ``render_contour`` :246-256, i.e.
OpenCV's ``drawContours(thickness=-1)`` polygon fill): contours of one image -> overlap-aware multi-channel label image.
"""
import numpy as np
import torch

from . import _lib as L

__all__ = ['contours2labels', 'resolve_label_channels', 'labels2property_table']

_channel_hint = {}   # (H, W) -> channels that sufficed in the previous contours2labels call


def contours2labels(contours, size, rounded=True, clip=True, initial_depth=1, gap=3, dtype='int32', ioa_thresh=None,
                    sort_by=None, sort_descending=True, return_indices=False, device=None):
    """Contours to labels (data/cpn.py:292-358).

    Args:
        contours: ``Array/Tensor[num_contours, num_points, 2]`` (x, y) of a single image (numpy or CUDA tensor).
        size: label image size ``(height, width)``.
        rounded, clip, initial_depth, gap, dtype: as in the reference.  ``sort_by`` / ``sort_descending`` reorder the
            contours first (labels then follow the sorted order, like the reference).
    Returns:
        ``[height, width, channels]`` label image: a CUDA ``int32`` tensor for tensor input, a numpy array of ``dtype``
        for numpy input.  Contour ``k`` carries label ``k + 1``; overlapping objects go to further channels.
    """
    if ioa_thresh is not None or return_indices:
        raise NotImplementedError('ioa_thresh / return_indices are outside the accelerated path.')
    as_numpy = not isinstance(contours, torch.Tensor)
    if as_numpy:
        if isinstance(contours, (list, tuple)):
            if len({np.asarray(c).shape for c in contours}) > 1:
                raise NotImplementedError('ragged contour lists are outside the accelerated path (pad or resample).')
            contours = np.stack([np.asarray(c) for c in contours]) if len(contours) else np.zeros((0, 1, 2), 'float32')
        dev = torch.device(device if device is not None else 'cuda')
        con = torch.as_tensor(np.asarray(contours, dtype=np.float32)).to(dev)
    else:
        if not contours.is_cuda:
            raise RuntimeError('celldetection_b200.data.contours2labels runs on CUDA tensors only (no CPU fallback).')
        con = contours.float()
    if sort_by is not None:
        order = torch.argsort(torch.as_tensor(sort_by).to(con.device), stable=True)
        if sort_descending:
            order = order.flip(0)
        con = con[order]
    con = con.contiguous()
    K = int(con.shape[0])
    S = int(con.shape[1]) if con.dim() == 3 else 1
    H, W = int(size[0]), int(size[1])
    lib = L.load()
    dev = con.device
    ws = torch.empty((int(lib.cpn_contours2labels_workspace_bytes(K, S)),), dtype=torch.uint8, device=dev)
    info = torch.zeros((4,), dtype=torch.int32, device=dev)
    # start from the channel count that sufficed last time for this image size (each retry re-zeroes the whole image)
    channels = max(int(initial_depth), 4, _channel_hint.get((H, W), 0))
    while True:
        labels = torch.empty((H, W, channels), dtype=torch.int32, device=dev)
        L.check(lib.cpn_contours2labels(L.ptr(con), K, S, H, W, int(bool(rounded)), int(bool(clip)), int(gap),
                                        L.ptr(labels), channels, L.ptr(ws), L.ptr(info), L.stream_ptr()),
                'contours2labels')
        used, needed, err, _ = info.tolist()
        if err:
            raise RuntimeError('contours2labels: dependency wait timed out (is the GPU shared with another process?)')
        if needed <= channels:
            break
        if channels >= 64:
            raise RuntimeError('contours2labels: more than 64 label channels needed')
        channels = min(64, max(needed, 2 * channels))
    _channel_hint[(H, W)] = channels if used > channels // 2 else max(4, channels // 2)
    labels = labels[:, :, :max(used, int(initial_depth))]
    if as_numpy:
        return labels.cpu().numpy().astype(dtype)
    return labels.contiguous()


def resolve_label_channels(labels, method='dilation', max_iter=999, kernel=(3, 3)):
    """Resolve label channels (data/cpn.py:361-398): ``[h, w, c]`` channel label image -> ``[h, w]`` flat label image.
    Pixels assigned to exactly one object keep its label; conflicts are filled from the neighbouring cores by repeated
    3x3-cross grey dilation.  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out."""
    if method != 'dilation':
        raise ValueError(f'Invalid method: {method}')
    if tuple(kernel) != (3, 3):
        raise NotImplementedError('only the default (3, 3) cross kernel is on the accelerated path.')
    as_numpy = not isinstance(labels, torch.Tensor)
    if as_numpy:
        dtype = np.asarray(labels).dtype
        lab = torch.as_tensor(np.ascontiguousarray(labels).astype(np.int32)).cuda()
    else:
        if not labels.is_cuda:
            raise RuntimeError('celldetection_b200.data.resolve_label_channels runs on CUDA tensors only.')
        lab = labels.to(torch.int32).contiguous()
    H, W, C = [int(v) for v in lab.shape]
    lib = L.load()
    flat = torch.empty((H, W), dtype=torch.int32, device=lab.device)
    ws = torch.empty((int(lib.cpn_resolve_label_channels_workspace_bytes(H, W)),), dtype=torch.uint8, device=lab.device)
    L.check(lib.cpn_resolve_label_channels(L.ptr(lab), H, W, C, int(max_iter), L.ptr(flat), L.ptr(ws), None,
                                           L.stream_ptr()), 'resolve_label_channels')
    return flat.cpu().numpy().astype(dtype) if as_numpy else flat


def labels2property_table(labels, *properties, **kwargs):
    """Labels to property table (data/misc.py:320-345); see ``utils.outputs.labels2property_table``."""
    from .utils.outputs import labels2property_table as impl
    return impl(labels, *properties, **kwargs)
