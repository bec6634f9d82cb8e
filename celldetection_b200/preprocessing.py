"""Input-side conditioning of a slide before tiling: the B200 counterpart of ``preprocess`` in
/root/reference/celldetection_scripts/cpn_inference.py:196-222 (percentile normalisation ``cd.data.normalize_percentile``,
data/misc.py:156-161; gamma; brightness / contrast; grayscale -> RGB replication).

The image goes to the GPU once; ``cpn_histogram`` builds its exact histogram, the host derives from those 256 / 65 536 counts
what the reference computes with whole-image numpy passes -- ``np.percentile``'s order statistics (same linear interpolation,
numpy's own lerp on the two neighbouring values), the image mean -- composes every step into ONE look-up table and
``cpn_apply_lut`` writes the uint8 image that ``apply_model`` then tiles as a device-resident slide.

Third-party pieces restated here (the packages are not part of this image, so these are pinned against their published
formulas only): ``skimage.img_as_ubyte`` for floats in [0, 1] = ``rint(x * 255)`` in float64; albumentations
``gamma_transform`` for uint8 = LUT ``((arange(256) / 255) ** gamma * 255).astype(uint8)``; ``brightness_contrast_adjust``
for uint8 = LUT ``clip(arange(256, float32) * alpha + beta * mean(img), 0, 255).astype(uint8)``.
"""
import numpy as np
import torch

from . import _lib as L

__all__ = ['preprocess', 'percentile_from_histogram', 'percentile_lut', 'tone_lut', 'build_lut']


def percentile_from_histogram(hist, percentiles):
    """``np.percentile(values, percentiles)`` (method 'linear') from the exact histogram of non-negative integers."""
    hist = np.asarray(hist, dtype=np.int64)
    n = int(hist.sum())
    assert n > 0, 'empty image'
    cdf = np.cumsum(hist)
    out = []
    for q in np.atleast_1d(np.asarray(percentiles, dtype=np.float64)):
        virtual = (n - 1) * np.true_divide(q, 100)
        prev = int(np.floor(virtual))
        nxt = min(prev + 1, n - 1)
        v0 = int(np.searchsorted(cdf, prev, side='right'))        # value of the order statistic with 0-based rank prev
        v1 = int(np.searchsorted(cdf, nxt, side='right'))
        t = np.float64(virtual - prev)
        a, b = np.float64(v0), np.float64(v1)
        # numpy's _lerp (lib/_function_base_impl.py) on the two neighbours, incl. its t >= 0.5 branch: bit-identical to the sort
        out.append(float(b - (b - a) * (1 - t)) if t >= 0.5 else float(a + (b - a) * t))
    return out


def percentile_lut(hist, bins, percentile=None):
    """``normalize_percentile`` (data/misc.py:156-161) + ``img_as_ubyte`` as a table over all ``bins`` input values."""
    pct = 99.9 if percentile is None else percentile                          # normalize_percentile's default
    if not isinstance(pct, (list, tuple)):
        pct = (100 - pct, pct)
    low, high = percentile_from_histogram(hist, pct)
    values = np.arange(bins, dtype=np.float64)
    x = (np.clip(values, low, high) - low) / (high - low)
    return np.clip(np.rint(x * 255.), 0, 255).astype(np.uint8)                # img_as_ubyte of a float image in [0, 1]


def tone_lut(hist8, gamma=1., contrast=1., brightness=0.):
    """Gamma, then contrast / brightness (which uses the mean of the image at that point) as one table over uint8 values;
    ``hist8``: histogram of the uint8 image the table is applied to.  ``None`` if the chain is the identity."""
    if gamma == 1. and contrast == 1.:
        return None
    lut = np.arange(256, dtype=np.uint8)
    if gamma != 1.:
        table = (np.arange(0, 256.0 / 255, 1.0 / 255) ** gamma) * 255
        lut = table.astype(np.uint8)[lut]
    if contrast != 1.:      # (the reference applies `brightness` only together with a contrast change, cpn_inference.py:220-221)
        h = np.bincount(lut, weights=np.asarray(hist8, dtype=np.float64), minlength=256)   # histogram after the gamma step
        mean = float((h * np.arange(256)).sum() / h.sum())
        t = np.arange(0, 256).astype(np.float32)
        t *= contrast
        if brightness != 0.:
            t += brightness * mean
        lut = np.clip(t, 0, 255).astype(np.uint8)[lut]
    return lut


def build_lut(hist, bins, percentile=None, gamma=1., contrast=1., brightness=0., implicit=False):
    """ONE uint8 table over all ``bins`` input values for the reference's chain without the grayscale branch:
    [percentile normalisation -> uint8] -> [gamma] -> [contrast / brightness]."""
    hist = np.asarray(hist, dtype=np.float64)
    if percentile is not None or implicit:
        lut = percentile_lut(hist, bins, percentile)
        hist8 = np.bincount(lut, weights=hist, minlength=256)
    else:
        assert bins == 256, 'images wider than 8 bit are percentile-normalised first (cpn_inference.py:200-202)'
        lut, hist8 = np.arange(256, dtype=np.uint8), hist
    tone = tone_lut(hist8, gamma, contrast, brightness)
    return lut if tone is None else tone[lut]


def _histogram(lib, t, dtype, bins):
    hist = torch.empty((bins,), dtype=torch.int32, device=t.device)
    L.check(lib.cpn_histogram(L.ptr(t), dtype, t.numel(), L.ptr(hist), L.stream_ptr()), 'histogram')
    return hist.cpu().numpy().view(np.uint32)


def _apply_lut(lib, t, dtype, lut):
    lut_d = torch.from_numpy(np.ascontiguousarray(lut)).to(t.device)
    out = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
    L.check(lib.cpn_apply_lut(L.ptr(t), dtype, t.numel(), L.ptr(lut_d), L.ptr(out), L.stream_ptr()), 'apply_lut')
    return out


def preprocess(img, gamma=1., contrast=1., brightness=0., percentile=None, grayscale=False, device='cuda'):
    """``Array[h, w(, c)]`` uint8 / uint16 (or a tensor of that type; ``torch.int16`` is read as the uint16 bit pattern, for
    torch builds whose ``from_numpy`` has no uint16) -> CUDA uint8 ``Tensor[h, w, 3 | c]``
    (cpn_inference.py:196-222; 2-D and single-channel images come back as three equal channels, :214-215)."""
    t = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img))
    if t.dtype == torch.uint8:
        dtype, bins = L.DT_U8, 256
    elif t.dtype in (torch.uint16, torch.int16):
        dtype, bins = L.DT_U16, 65536
    else:
        raise NotImplementedError(f'preprocess: {t.dtype} images are not supported on the accelerated path (uint8 / uint16 are)')
    t = t.to(device).contiguous()
    if t.dim() == 2:
        t = t[..., None]
    C = int(t.shape[-1])
    implicit = bins > 256                                   # "implicit percentile normalization, since input is not uint8"
    normalise = percentile is not None or implicit
    to_gray = grayscale and C > 1
    if to_gray and C not in (3, 4):
        raise NotImplementedError('grayscale: 3 (RGB) or 4 (RGBA) channels (2-channel mean is outside the accelerated path)')
    lib = L.load()
    if not to_gray:
        if not normalise and gamma == 1. and contrast == 1.:
            out = t
        else:
            lut = build_lut(_histogram(lib, t, dtype, bins), bins, percentile, gamma, contrast, brightness, implicit)
            out = _apply_lut(lib, t, dtype, lut)
    else:
        # the channel mix sits between the two tables: (percentile table ->) gray -> histogram of the gray image -> tone table
        lut1 = percentile_lut(_histogram(lib, t, dtype, bins), bins, percentile) if normalise else None
        lut1_d = None if lut1 is None else torch.from_numpy(lut1).to(t.device)
        out = torch.empty(t.shape[:2] + (1,), dtype=torch.uint8, device=t.device)
        L.check(lib.cpn_rgb2gray(L.ptr(t), dtype, out.numel(), C, L.ptr(lut1_d), L.ptr(out), L.stream_ptr()), 'rgb2gray')
        if gamma != 1. or contrast != 1.:
            out = _apply_lut(lib, out, L.DT_U8, tone_lut(_histogram(lib, out, L.DT_U8, 256), gamma, contrast, brightness))
    if out.shape[-1] == 1:
        out = out.expand(-1, -1, 3).contiguous()            # cv2.COLOR_GRAY2RGB (cpn_inference.py:214-215)
    return out
