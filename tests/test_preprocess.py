"""Slide preprocessing (cpn_inference.py:196-222): CPU tests pin the host logic (np.percentile from a histogram, the composed
look-up tables) and the oracle against the reference-minted golden; GPU tests run ``cd.preprocess`` through the C ABI
(cpn_histogram / cpn_apply_lut / cpn_rgb2gray) against the same vectors and against the oracle on larger images."""
import numpy as np
import pytest
import torch

import celldetection_b200 as cd
from celldetection_b200 import preprocessing as P
from helpers import load_npz

import preprocess_oracle as po


def _cases(z):
    for name in ('u16', 'u8', 'rgba'):
        j = 0
        while f'{name}/chain{j}/params' in z.files:
            gamma, con, bri, pct, gray = z[f'{name}/chain{j}/params']
            yield name, z[f'{name}/img'], dict(gamma=float(gamma), contrast=float(con), brightness=float(bri),
                                                percentile=None if pct < 0 else float(pct),
                                                grayscale=bool(gray)), z[f'{name}/chain{j}/out']
            j += 1


def test_oracle_reproduces_golden():
    z = load_npz('preprocess')
    n = 0
    for name, img, kw, want in _cases(z):
        assert np.array_equal(po.preprocess(img, **kw), want), (name, kw)
        n += 1
    assert n == 11
    for name in ('u16', 'u8'):
        for j in range(3):
            pct = z[f'{name}/pct{j}'].tolist()
            pct = pct[0] if len(pct) == 1 else pct
            assert np.array_equal(po.normalize_percentile(z[f'{name}/img'], pct), z[f'{name}/norm{j}'])


def test_rgb2gray_formula_equals_cv2_on_every_colour():
    cv2 = pytest.importorskip('cv2')
    v = np.arange(256, dtype=np.uint8)
    img = np.stack(np.meshgrid(v, v, v, indexing='ij'), -1).reshape(4096, 4096, 3)
    assert np.array_equal(po.rgb2gray(img), cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))


@pytest.mark.parametrize('seed', range(6))
def test_percentile_from_histogram_is_numpys(seed):
    """The order statistics and numpy's own interpolation from the exact histogram == np.percentile on the values."""
    rng = np.random.RandomState(seed)
    n = int(rng.randint(1, 5000))
    bins = 256 if seed % 2 else 65536
    vals = (rng.gamma(2., bins / 40., size=n)).clip(0, bins - 1).astype(np.uint16)
    if seed == 4:
        vals[:] = 7                                                   # constant image
    hist = np.bincount(vals, minlength=bins)
    qs = [0., 0.1, 0.5, 2., 50., 98., 99.5, 99.9, 100.]
    assert P.percentile_from_histogram(hist, qs) == [float(v) for v in np.percentile(vals, qs)]


def test_host_tables_reproduce_golden():
    """Non-grayscale chains are ONE table: lut[img] must be the reference chain's output."""
    z = load_npz('preprocess')
    for name, img, kw, want in _cases(z):
        if kw.pop('grayscale'):
            continue
        bins = 256 if img.dtype == np.uint8 else 65536
        hist = np.bincount(img.reshape(-1), minlength=bins)
        lut = P.build_lut(hist, bins, implicit=bins > 256, **kw)
        got = lut[img]
        if got.ndim == 2:
            got = np.repeat(got[..., None], 3, -1)
        assert np.array_equal(got, want), (name, kw)


@pytest.mark.gpu
def test_preprocess_matches_golden_on_gpu():
    z = load_npz('preprocess')
    for name, img, kw, want in _cases(z):
        src = img.view(np.int16) if img.dtype == np.uint16 else img        # torch.from_numpy has no uint16 everywhere
        t = torch.from_numpy(np.ascontiguousarray(src))
        got = cd.preprocess(t, **kw).cpu().numpy()
        assert got.dtype == np.uint8 and np.array_equal(got, want), (name, kw)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype,shape', [(np.uint8, (1031, 777, 3)), (np.uint16, (2049, 1023)), (np.uint16, (515, 640, 3)),
                                         (np.uint8, (3, 5)), (np.uint8, (600, 700, 4))])
def test_preprocess_equals_oracle_on_larger_images(dtype, shape):
    """Ragged sizes (vector body + scalar tail), 2-D and multi-channel inputs, all option combinations."""
    rng = np.random.RandomState(len(shape) * 7 + shape[0])
    top = 255 if dtype == np.uint8 else 65535
    img = rng.gamma(2., top / 25., size=shape).clip(0, top).astype(dtype)
    for kw in (dict(percentile=99.), dict(gamma=0.7), dict(contrast=1.4, brightness=-0.1), dict(),
               dict(percentile=[1., 99.5], gamma=1.3, contrast=0.8, brightness=0.2),
               dict(grayscale=True, gamma=1.1, contrast=1.2), dict(grayscale=True, percentile=98.5)):
        if kw.get('grayscale') and img.ndim == 3 and img.shape[-1] == 4 and dtype != np.uint8:
            continue
        want = po.preprocess(img, **kw)
        if want.ndim == 3 and want.shape[-1] == 4 and not kw.get('grayscale'):
            pass                                                           # RGBA without grayscale stays RGBA
        src = img.view(np.int16) if dtype == np.uint16 else img
        got = cd.preprocess(torch.from_numpy(np.ascontiguousarray(src)), **kw).cpu().numpy()
        assert np.array_equal(got, want), (dtype, shape, kw)


@pytest.mark.gpu
def test_histogram_and_lut_entry_points():
    """cpn_histogram is exact for both widths (incl. unaligned bases and empty input); cpn_apply_lut is plain indexing."""
    from celldetection_b200 import _lib as L
    lib = L.load()
    rng = np.random.RandomState(0)
    for dt, code, bins in ((np.uint8, L.DT_U8, 256), (np.uint16, L.DT_U16, 65536)):
        for n, off in ((0, 0), (1, 0), (4097, 0), (100003, 1), (1 << 20, 0)):
            vals = rng.randint(0, bins, size=n + off).astype(dt)
            if n > 5000:
                vals[::3] = 11                                             # a dominant bin (merged atomics)
            src = vals.view(np.int16) if dt == np.uint16 else vals
            d = torch.from_numpy(src).cuda()[off:]
            hist = torch.empty((bins,), dtype=torch.int32, device='cuda')
            L.check(lib.cpn_histogram(L.ptr(d), code, n, L.ptr(hist), L.stream_ptr()), 'histogram')
            assert np.array_equal(hist.cpu().numpy(), np.bincount(vals[off:], minlength=bins))
            lut = rng.randint(0, 256, size=bins).astype(np.uint8)
            out = torch.empty((n,), dtype=torch.uint8, device='cuda')
            L.check(lib.cpn_apply_lut(L.ptr(d), code, n, L.ptr(torch.from_numpy(lut).cuda()), L.ptr(out), L.stream_ptr()), 'lut')
            assert np.array_equal(out.cpu().numpy(), lut[vals[off:]])
