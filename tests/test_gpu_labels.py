"""GPU: cd.data.contours2labels (cpn_contours2labels) against label images minted from the reference's own
contours2labels (data/cpn.py:292-358 + cv2.drawContours) and against the oracle at larger sizes.  Integer work: bit-exact."""
import numpy as np
import pytest
import torch

import celldetection_b200 as cd
import c2l_oracle as c2l
from helpers import load_npz

pytestmark = pytest.mark.gpu


def synth_contours(rng, K, S, H, W, radius=(2., 14.)):
    t = np.linspace(0, 1, S)
    out = np.zeros((K, S, 2), np.float32)
    for k in range(K):
        order, r = rng.randint(1, 6), rng.uniform(*radius)
        xy = np.array([rng.rand() * W, rng.rand() * H])[None] + sum(
            rng.randn(2)[None] * r / (j + 1) * np.cos(2 * np.pi * (j + 1) * t)[:, None] +
            rng.randn(2)[None] * r / (j + 1) * np.sin(2 * np.pi * (j + 1) * t)[:, None] for j in range(order))
        out[k] = xy
    return out


@pytest.mark.parametrize('name', ['sparse', 'dense', 'border', 'odd'])
def test_contours2labels_matches_reference_golden(name):
    z = load_npz('contours2labels')
    H, W = [int(v) for v in z[f'{name}/size']]
    want = z[f'{name}/labels']
    got = cd.data.contours2labels(z[f'{name}/contours'], (H, W))                 # numpy in -> numpy out
    assert got.dtype == np.int32 and got.shape == want.shape
    assert np.array_equal(got, want)
    got_t = cd.data.contours2labels(torch.from_numpy(z[f'{name}/contours']).cuda(), (H, W))   # tensor in -> CUDA tensor
    assert got_t.is_cuda and got_t.dtype == torch.int32 and np.array_equal(got_t.cpu().numpy(), want)


def test_contours2labels_variants_and_edge_cases():
    z = load_npz('contours2labels')
    con = z['variants/contours']
    for tag, kw in (('noround', dict(rounded=False)), ('gap0', dict(gap=0)), ('depth3', dict(initial_depth=3))):
        got = cd.data.contours2labels(con, (80, 80), **kw)
        assert got.shape == z[f'{tag}/labels'].shape and np.array_equal(got, z[f'{tag}/labels']), tag
    empty = cd.data.contours2labels(np.zeros((0, 32, 2), np.float32), (40, 50))
    assert empty.shape == (40, 50, 1) and not empty.any()
    # sort_by: labels follow the sorted order (data/cpn.py:331-335)
    scores = np.random.RandomState(1).rand(len(con))
    got = cd.data.contours2labels(con, (80, 80), sort_by=scores)
    want = c2l.contours2labels(con[np.argsort(scores)[::-1]].copy(), (80, 80))
    assert np.array_equal(got, want)
    with pytest.raises(NotImplementedError):
        cd.data.contours2labels(con, (80, 80), ioa_thresh=0.5)
    with pytest.raises(RuntimeError):
        cd.data.contours2labels(torch.from_numpy(con), (80, 80))                 # CPU tensor: no fallback


def test_contours2labels_medium_against_oracle_and_determinism():
    """4000 contours on 1536 x 2048 (deep dependency chains across the whole persistent grid): equal to the oracle bit
    for bit, and identical across repeated runs (the ordered schedule is deterministic)."""
    rng = np.random.RandomState(7)
    H, W, K, S = 1536, 2048, 4000, 32
    con = synth_contours(rng, K, S, H, W, (4., 22.))
    want = c2l.contours2labels(con.copy(), (H, W))
    t = torch.from_numpy(con).cuda()
    got = cd.data.contours2labels(t, (H, W))
    assert tuple(got.shape) == want.shape and np.array_equal(got.cpu().numpy(), want)
    again = cd.data.contours2labels(t, (H, W))
    assert torch.equal(got, again)
    # every label is present exactly where the oracle painted it: label k + 1 <-> contour k
    ids = torch.unique(got)
    assert int(ids.max()) <= K and int(ids.min()) == 0


def test_contours2labels_full_size_properties():
    """Config-C4-sized label image (16384 x 16384, 1e5 contours of 128 samples): size-independent properties --
    every contour's label appears, only inside its own bounding box, in exactly one channel; channel 0 region rule holds
    for a sample of contours (no other label of the same channel inside the gap-dilated box)."""
    rng = np.random.RandomState(11)
    H = W = 16384
    K, S = 100000, 128
    base = synth_contours(rng, 500, S, 64, 64, (4., 20.)) - 32.
    centres = np.stack((rng.rand(K) * W, rng.rand(K) * H), -1).astype(np.float32)
    con = base[rng.randint(0, 500, K)] + centres[:, None]
    t = torch.from_numpy(con.astype(np.float32)).cuda()
    lab = cd.data.contours2labels(t, (H, W))
    C = lab.shape[2]
    assert 1 <= C <= 64
    flat = lab.reshape(-1, C)
    counts = torch.bincount(flat.reshape(-1).long(), minlength=K + 1)
    assert int((counts[1:] > 0).sum()) == K                       # every contour painted
    r = torch.round(t)
    r[..., 0].clamp_(0, W - 1), r[..., 1].clamp_(0, H - 1)
    mn, mx = r.min(1).values.long().cpu().numpy(), r.max(1).values.long().cpu().numpy()
    for k in rng.randint(0, K, 200):
        x0, y0, x1, y1 = mn[k, 0], mn[k, 1], mx[k, 0], mx[k, 1]
        crop = lab[max(y0 - 3, 0):y1 + 4, max(x0 - 3, 0):x1 + 4]
        ch = [c for c in range(C) if bool((crop[..., c] == k + 1).any())]
        assert len(ch) == 1
        vals = torch.unique(crop[..., ch[0]])
        earlier = [int(v) for v in vals if 0 < int(v) < k + 1]
        assert not earlier, (k, earlier)                          # no earlier label inside the dilated box of k's channel
        assert int(counts[k + 1]) == int((crop[..., ch[0]] == k + 1).sum())   # painted only inside its own box


@pytest.mark.parametrize('name', ['sparse', 'dense', 'border', 'odd'])
def test_resolve_label_channels_matches_reference_golden(name):
    """cd.data.resolve_label_channels (data/cpn.py:361-398) on the reference's own label images: bit-exact."""
    z = load_npz('contours2labels')
    lab = z[f'{name}/labels'].astype(np.int32)
    got = cd.data.resolve_label_channels(lab)
    assert got.shape == z[f'{name}/flat'].shape and got.dtype == np.int32
    assert np.array_equal(got, z[f'{name}/flat'])
    t = cd.data.resolve_label_channels(torch.from_numpy(lab).cuda())
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), z[f'{name}/flat'])
    one = lab[..., :1]                                                          # no overlap anywhere -> labels.max(-1)
    assert np.array_equal(cd.data.resolve_label_channels(one), one[..., 0])
    neg = lab.copy()
    neg[0, 0, :] = -1                                                           # an ignore label on a background pixel
    assert np.array_equal(cd.data.resolve_label_channels(neg), c2l.resolve_label_channels(neg))
    with pytest.raises(NotImplementedError):
        cd.data.resolve_label_channels(lab, kernel=(5, 5))


def test_resolve_label_channels_medium_against_oracle():
    rng = np.random.RandomState(3)
    H, W = 1024, 1536
    con = synth_contours(rng, 2500, 32, H, W, (5., 30.))
    lab = cd.data.contours2labels(torch.from_numpy(con).cuda(), (H, W))
    flat = cd.data.resolve_label_channels(lab)
    want = c2l.resolve_label_channels(lab.cpu().numpy())
    assert np.array_equal(flat.cpu().numpy(), want)
    # every overlap pixel got a label of one of the objects covering its neighbourhood; cores are untouched
    cnt = (lab > 0).sum(-1)
    assert torch.equal(flat[cnt == 1], lab.max(-1).values[cnt == 1]) and int((flat[cnt == 0] != 0).sum()) == 0
