"""CPU: the oracle restatement against the golden vectors minted from the reference itself (oracle/make_golden.py).

This is what pins the oracle (the reference ships no golden vectors for this path)."""
import numpy as np
import pytest
import torch

import cpn_oracle as orc
from helpers import (load_npz, fixture_state_dict, fixture_ctor, ensemble_state_dicts, rel_err, MODEL_FIXTURES,
                     VARIANT_FIXTURES)


@pytest.mark.parametrize('name', MODEL_FIXTURES)
def test_oracle_model_matches_reference_golden(name):
    z = load_npz(name)
    arch = str(z['arch'])
    n, h, w, seed, order, samples = [int(v) for v in z['meta']]
    sd = fixture_state_dict(z, arch, seed)
    x = torch.from_numpy(z['x'])
    torch.set_num_threads(8)
    with torch.no_grad():
        scores, locations, refinement, fourier = orc.cpn_core(x, sd, arch)
    # fp32 torch-CPU restatement of the same ops: agreement to rounding level
    assert rel_err(scores, z['raw_scores']) < 1e-5
    assert rel_err(locations, z['raw_locations']) < 1e-5
    assert rel_err(refinement, z['raw_refinement']) < 1e-5
    assert rel_err(fourier, z['raw_fourier']) < 1e-5
    offsets = torch.from_numpy(z['offsets']) if 'offsets' in z.files else None
    # post chain on the *reference's* head tensors: isolates it from conv rounding -> exact counts
    out = orc.cpn_post(torch.from_numpy(z['raw_scores']), torch.from_numpy(z['raw_locations']),
                       torch.from_numpy(z['raw_refinement']), torch.from_numpy(z['raw_fourier']), (h, w), order=order,
                       samples=samples, offsets=offsets)
    for i in range(n):
        assert len(out['scores'][i]) == len(z[f'out/{i}/scores'])
        for k in ('contours', 'boxes', 'scores', 'locations', 'fourier', 'contour_proposals'):
            assert rel_err(out[k][i].numpy(), z[f'out/{i}/{k}']) < 1e-6, (k, i)
        assert np.array_equal(out['classes'][i].numpy(), z[f'out/{i}/classes'])
    nonms = orc.cpn_post(torch.from_numpy(z['raw_scores']), torch.from_numpy(z['raw_locations']),
                         torch.from_numpy(z['raw_refinement']), torch.from_numpy(z['raw_fourier']), (h, w), order=order,
                         samples=samples, nms_on=False)
    for i in range(n):
        assert len(nonms['scores'][i]) == int(z[f'nonms_count/{i}'])


@pytest.mark.parametrize('name', VARIANT_FIXTURES)
def test_oracle_variant_models_match_reference_golden(name):
    """classes > 2 (softmax / argmax), uncertainty head (certainty filter, uncertainty_nms, box_uncertainties) and
    bucketed refinement: oracle vs vectors minted from the reference (models/cpn.py:73-82, 583-585, 617-618, 723-726)."""
    z = load_npz(name)
    arch = str(z['arch'])
    n, h, w, seed, order, samples = [int(v) for v in z['meta']]
    ctor, attrs = fixture_ctor(z)
    sd = fixture_state_dict(z, arch, seed)
    x = torch.from_numpy(z['x'])
    torch.set_num_threads(8)
    with torch.no_grad():
        raw = orc.cpn_core5(x, sd, arch, **{k: ctor[k] for k in orc.CORE_KW if k in ctor})
    for nm, t in zip(('scores', 'locations', 'refinement', 'fourier', 'uncertainty'), raw):
        if t is not None:
            assert rel_err(t, z['raw_' + nm]) < 1e-5, nm
    assert (raw[4] is not None) == bool(ctor.get('uncertainty_head'))
    assert raw[2].shape[1] == 2 * ctor.get('refinement_buckets', 1)
    unc = torch.from_numpy(z['raw_uncertainty']) if 'raw_uncertainty' in z.files else None
    offsets = torch.from_numpy(z['offsets']) if 'offsets' in z.files else None
    out = orc.cpn_post(torch.from_numpy(z['raw_scores']), torch.from_numpy(z['raw_locations']),
                       torch.from_numpy(z['raw_refinement']), torch.from_numpy(z['raw_fourier']), (h, w), order=order,
                       samples=samples, offsets=offsets, uncertainty=unc, **attrs)
    keys = ['contours', 'boxes', 'scores', 'locations', 'fourier', 'contour_proposals']
    if unc is not None:
        keys.append('box_uncertainties')
    for i in range(n):
        assert len(out['scores'][i]) == len(z[f'out/{i}/scores']) > 0
        for k in keys:
            assert rel_err(out[k][i].numpy(), z[f'out/{i}/{k}']) < 1e-6, (k, i)
        assert np.array_equal(out['classes'][i].numpy(), z[f'out/{i}/classes'])


def test_oracle_fouriers2contours_golden():
    z = load_npz('fouriers2contours')
    for order in (1, 5, 16):
        for samples in (32, 64, 128):
            t = f'o{order}_s{samples}'
            con, _ = orc.fouriers2contours(torch.from_numpy(z[t + '/fourier']), torch.from_numpy(z[t + '/locations']),
                                           samples=samples)
            assert np.abs(con.numpy() - z[t + '/contours']).max() < 1e-4  # px
    con, _ = orc.fouriers2contours(torch.from_numpy(z['explicit/fourier']), torch.from_numpy(z['explicit/locations']),
                                   sampling=torch.from_numpy(z['explicit/sampling']))
    assert np.abs(con.numpy() - z['explicit/contours']).max() < 1e-4


def test_oracle_tiling_golden():
    z = load_npz('tiling')
    i = 0
    while f'{i}/args' in z.files:
        size, crop, strides = [tuple(int(v) for v in r) for r in z[f'{i}/args']]
        sl, ov, shape = orc.get_tiling_slices(size, crop, strides)
        assert np.array_equal(np.array([[[s.start, s.stop] for s in t] for t in sl]), z[f'{i}/slices'])
        assert np.array_equal(ov, z[f'{i}/overlaps'])
        assert tuple(shape) == tuple(z[f'{i}/shape'])
        i += 1
    assert i == 5
    # config C4: 16384^2, crop 512 -> 43 x 43 tiles at stride 384, 32 x 32 at stride 512
    assert orc.get_tiling_slices((16384, 16384), (512, 512), (384, 384))[2] == (43, 43)
    assert orc.get_tiling_slices((16384, 16384), (512, 512), (512, 512))[2] == (32, 32)


def test_oracle_apply_model_golden():
    z = load_npz('apply_model_cpnu22')
    seed, crop, stride, border = [int(v) for v in z['meta']]
    sd = fixture_state_dict(z, 'CpnU22', seed)
    torch.set_num_threads(8)
    res = orc.apply_model(z['img'], sd, 'CpnU22', crop, stride, border_removal=border)
    assert len(res['scores']) == len(z['out/scores']) > 0
    for k in ('contours', 'boxes', 'scores', 'locations', 'fourier', 'contour_proposals'):
        assert rel_err(res[k].numpy(), z['out/' + k]) < 1e-5, k


def test_oracle_apply_model_test_time_repetitions_golden():
    """reps=3 + a host transform of the crop (TileLoader, cpn_inference.py:85-91,114-118), minted by the reference."""
    z = load_npz('apply_model_cpnu22')
    seed, crop, stride, border = [int(v) for v in z['meta']]
    sd = fixture_state_dict(z, 'CpnU22', seed)
    torch.set_num_threads(8)
    res = orc.apply_model(z['img'], sd, 'CpnU22', crop, stride, border_removal=border, reps=3,
                          transforms=orc.tta_example_transform)
    assert len(res['scores']) == len(z['tta/scores']) > len(z['out/scores'])
    for k in ('contours', 'boxes', 'scores', 'locations', 'fourier', 'contour_proposals'):
        assert rel_err(res[k].numpy(), z['tta/' + k]) < 1e-5, k


def test_oracle_nms_matches_torchvision():
    """The NMS restatement against the third-party op the reference calls (torchvision)."""
    import torchvision  # noqa: F401
    g = torch.Generator().manual_seed(0)
    for n in (0, 1, 2, 50, 700):
        xy = torch.rand(n, 2, generator=g) * 100
        wh = torch.rand(n, 2, generator=g) * 30
        boxes = torch.cat((xy, xy + wh), 1)
        scores = torch.rand(n, generator=g)
        if n >= 50:  # ties and degenerate boxes
            scores[10:20] = scores[10]
            boxes[5] = torch.tensor([5., 5., 5., 5.])
            boxes[6] = torch.tensor([5., 5., 5., 5.])
        for thr in (0.2, 0.5):
            ref = torch.ops.torchvision.nms(boxes, scores, thr).numpy()
            got = orc.nms(boxes.numpy(), scores.numpy(), thr)
            assert np.array_equal(ref, got), (n, thr)
    # appendix A.3 cases
    z = torch.tensor([[5., 5., 5., 5.], [5., 5., 5., 5.]])
    assert list(orc.nms(z.numpy(), np.array([.5, .4], np.float32), .2)) == [0, 1]
    b = torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 10.]])
    assert list(orc.nms(b.numpy(), np.array([.5, .5], np.float32), 1.0)) == [0, 1]   # IoU == thr keeps both
    assert list(orc.nms(b.numpy(), np.array([.5, .5], np.float32), .99)) == [0]


def test_oracle_chunked_nms_rule():
    g = torch.Generator().manual_seed(1)
    n = 300
    xy = torch.rand(n, 2, generator=g) * 60
    boxes = torch.cat((xy, xy + torch.rand(n, 2, generator=g) * 20 + 1), 1)
    scores = torch.rand(n, generator=g)
    # reference formulation (ops/cpn.py:213-224) spelled out with torchvision's op
    idx = torch.zeros(0, dtype=torch.long)
    for s in range(0, n, 128):
        e = min(s + 128, n)
        idx = torch.cat((idx, torch.ops.torchvision.nms(boxes[s:e], scores[s:e], .3) + s))
    idx = idx[torch.ops.torchvision.nms(boxes[idx], scores[idx], .3)]
    got = orc.batched_box_nmsi([boxes.numpy()], [scores.numpy()], .3, batch_size=128)[0]
    assert np.array_equal(got, idx.numpy())


def test_oracle_ensemble_mask_and_voting_golden():
    """Two-model ensemble with a mask (upper score bound, empty tiles skipped) through the reference's apply_model, and
    cd.ops.filter_by_box_voting at op level (cpn_inference.py:93-111, 417-427; ops/boxes.py:53-83).  With min_vote > 1
    the reference's driver raises IndexError whenever the vote removes a box (votes are stored before the keep
    indices are applied, cpn_inference.py:421-423), so the driver-level vector uses min_vote = 1."""
    z = load_npz('apply_model_ensemble')
    crop, stride, border = [int(v) for v in z['meta']]
    sds = ensemble_state_dicts(z)
    torch.set_num_threads(8)
    res = orc.apply_models(z['img'], sds, ['CpnU22'] * 2, crop, stride, border_removal=border, mask=z['mask'], min_vote=1)
    assert len(res['scores']) == len(z['vote1/scores']) > 0
    for k in ('contours', 'boxes', 'scores', 'locations', 'fourier', 'contour_proposals'):
        assert rel_err(res[k].numpy(), z['vote1/' + k]) < 1e-5, k
    # every detection lies inside the mask's bounding region (the mask is the upper score bound)
    cx = res['locations'][:, 0].numpy()
    assert cx.max() < 125
    keep, votes = orc.filter_by_box_voting(z['voting/boxes'], 0.2, 2)
    assert np.array_equal(keep.numpy(), z['voting/keep'])
    assert np.abs(votes.numpy() - z['voting/votes']).max() < 1e-5
    # the sane order (filter, then attach votes) works where the reference crashes
    res2 = orc.apply_models(z['img'], sds, ['CpnU22'] * 2, crop, stride, border_removal=border, mask=z['mask'],
                            min_vote=1.5)
    assert 0 < len(res2['scores']) <= len(res['scores']) and len(res2['votes']) == len(res2['scores'])
    assert float(res2['votes'].min()) >= 1.5


C2L_CASES = ['sparse', 'dense', 'border', 'odd']


def test_oracle_contours2labels_golden():
    """Label rasterisation (data/cpn.py:292-358): the oracle's restated OpenCV polygon fill + greedy channel rule against
    label images minted from the reference's own contours2labels."""
    import c2l_oracle as c2l
    z = load_npz('contours2labels')
    for name in C2L_CASES:
        H, W = [int(v) for v in z[f'{name}/size']]
        got = c2l.contours2labels(z[f'{name}/contours'].copy(), (H, W))
        assert got.shape == z[f'{name}/labels'].shape and np.array_equal(got, z[f'{name}/labels']), name
        flat = c2l.resolve_label_channels(got)                                   # data/cpn.py:361-398
        assert np.array_equal(flat, z[f'{name}/flat']), name
    con = z['variants/contours']
    for tag, kw in (('noround', dict(rounded=False)), ('gap0', dict(gap=0)), ('depth3', dict(initial_depth=3))):
        got = c2l.contours2labels(con.copy(), (80, 80), **kw)
        assert got.shape == z[f'{tag}/labels'].shape and np.array_equal(got, z[f'{tag}/labels']), tag


def test_oracle_polygon_fill_matches_opencv():
    """The third-party piece of the path: cv2.drawContours(thickness=-1) (OpenCV, unpinned by the reference; the
    installed version is the oracle) against the restated Bresenham + scan-line fill, on random integer polygons."""
    cv2 = pytest.importorskip('cv2')
    import c2l_oracle as c2l
    rng = np.random.RandomState(0)
    for trial in range(400):
        n = rng.randint(1, 40)
        pts = rng.randint(0, [12, 40, 25][trial % 3], size=(n, 2)).astype(np.int32)
        xmin, ymin = pts.min(0)
        xmax, ymax = pts.max(0)
        a = np.zeros((ymax - ymin + 1, xmax - xmin + 1), np.int32)
        a = cv2.drawContours(a, [pts.reshape(-1, 1, 2)], 0, 1, -1, offset=(int(-xmin), int(-ymin)))
        m = c2l.fill_polygon(pts - [xmin, ymin], a.shape[1], a.shape[0])
        assert np.array_equal(a > 0, m), pts.tolist()
