"""GPU: the post-head chain, fouriers2contours, NMS and border filter through the C ABI against the oracle / golden
vectors / torchvision's CPU op.  Tolerances are stated per assertion (bit-exact for indices and counts)."""
import numpy as np
import pytest
import torch

import celldetection_b200 as cd
import cpn_oracle as orc
from helpers import load_npz, rel_err, fixture_ctor, MODEL_FIXTURES, VARIANT_FIXTURES

pytestmark = pytest.mark.gpu


def _post_model(z, **kw):
    n, h, w, seed, order, samples = [int(v) for v in z['meta']]
    m = cd.models.CPN(str(z['arch']), order=order, samples=samples, **kw).cuda()
    return m, (n, h, w, order, samples)


@pytest.mark.parametrize('name', MODEL_FIXTURES)
def test_post_chain_on_reference_head_tensors(name):
    """Select / decode / refine / boxes / NMS on the *reference's* head tensors: identical selection and keep sets,
    values within 1e-4 px (expected: bit-exact up to the last ulp of sigmoid)."""
    z = load_npz(name)
    m, (n, h, w, order, samples) = _post_model(z)
    scores = torch.from_numpy(z['raw_scores'])[:, 0].contiguous().cuda()
    locfou = torch.cat((torch.from_numpy(z['raw_locations']), torch.from_numpy(z['raw_fourier'])), 1)
    locfou = locfou.permute(0, 2, 3, 1).contiguous().cuda()
    ref = torch.from_numpy(z['raw_refinement']).permute(0, 2, 3, 1).contiguous().cuda()
    offsets = torch.from_numpy(z['offsets']).cuda() if 'offsets' in z.files else None
    out = m.post(scores, locfou, ref, (h, w), offsets=offsets)
    nonms = m.post(scores, locfou, ref, (h, w), nms=False)
    for i in range(n):
        assert len(nonms['scores'][i]) == int(z[f'nonms_count/{i}'])
        assert len(out['scores'][i]) == len(z[f'out/{i}/scores'])
        for k in ('contours', 'boxes', 'locations', 'fourier', 'contour_proposals'):
            assert np.abs(out[k][i].cpu().numpy() - z[f'out/{i}/{k}']).max() < 1e-4, (k, i)
        assert np.abs(out['scores'][i].cpu().numpy() - z[f'out/{i}/scores']).max() < 1e-6
        assert np.array_equal(out['classes'][i].cpu().numpy(), z[f'out/{i}/classes'])
    assert out['box_uncertainties'] is None


@pytest.mark.parametrize('name', VARIANT_FIXTURES)
def test_post_chain_variants_on_reference_head_tensors(name):
    """Softmax / argmax selection, certainty filter, uncertainty-weighted NMS and bucketed refinement on the
    *reference's* head tensors: identical selection / keep sets / classes, values within 1e-4 px."""
    z = load_npz(name)
    ctor, attrs = fixture_ctor(z)
    m, (n, h, w, order, samples) = _post_model(z, **ctor)
    for k, v in attrs.items():
        setattr(m, k, v)
    rs = torch.from_numpy(z['raw_scores'])
    scores = (rs[:, 0] if rs.shape[1] == 1 else rs.permute(0, 2, 3, 1)).contiguous().cuda()
    locfou = torch.cat((torch.from_numpy(z['raw_locations']), torch.from_numpy(z['raw_fourier'])), 1)
    locfou = locfou.permute(0, 2, 3, 1).contiguous().cuda()
    ref = torch.from_numpy(z['raw_refinement']).permute(0, 2, 3, 1).contiguous().cuda()
    unc = None
    if 'raw_uncertainty' in z.files:
        unc = torch.from_numpy(z['raw_uncertainty']).permute(0, 2, 3, 1).contiguous().cuda()
    offsets = torch.from_numpy(z['offsets']).cuda() if 'offsets' in z.files else None
    out = m.post(scores, locfou, ref, (h, w), offsets=offsets, uncertainty=unc)
    nonms = m.post(scores, locfou, ref, (h, w), nms=False, uncertainty=unc)
    keys = ['contours', 'boxes', 'locations', 'fourier', 'contour_proposals']
    for i in range(n):
        assert len(nonms['scores'][i]) == int(z[f'nonms_count/{i}'])
        assert len(out['scores'][i]) == len(z[f'out/{i}/scores']) > 0
        for k in keys:
            assert np.abs(out[k][i].cpu().numpy() - z[f'out/{i}/{k}']).max() < 1e-4, (k, i)
        assert np.abs(out['scores'][i].cpu().numpy() - z[f'out/{i}/scores']).max() < 1e-6
        assert np.array_equal(out['classes'][i].cpu().numpy(), z[f'out/{i}/classes'])
        if unc is not None:
            assert np.abs(out['box_uncertainties'][i].cpu().numpy() - z[f'out/{i}/box_uncertainties']).max() < 1e-7
    if unc is None:
        assert out['box_uncertainties'] is None
    # the certainty filter is a runtime attribute (cpn.py:617-618): turning it off can only add proposals
    if unc is not None and m.certainty_thresh is not None:
        m.certainty_thresh = None
        more = m.post(scores, locfou, ref, (h, w), nms=False, uncertainty=unc)
        assert all(len(a) >= len(b) for a, b in zip(more['scores'], nonms['scores']))
        assert sum(len(a) for a in more['scores']) > sum(len(b) for b in nonms['scores'])


def test_post_chain_runtime_attributes_and_edge_cases():
    z = load_npz(MODEL_FIXTURES[0])
    m, (n, h, w, order, samples) = _post_model(z)
    raw = [torch.from_numpy(z[k]) for k in ('raw_scores', 'raw_locations', 'raw_refinement', 'raw_fourier')]
    scores = raw[0][:, 0].contiguous().cuda()
    locfou = torch.cat((raw[1], raw[3]), 1).permute(0, 2, 3, 1).contiguous().cuda()
    ref = raw[2].permute(0, 2, 3, 1).contiguous().cuda()
    for kw in (dict(samples=64), dict(order=3), dict(refinement_iterations=0), dict(score_thresh=0.97),
               dict(nms_thresh=0.6), dict(samples=100, order=2)):
        okw = dict(order=order, samples=samples)
        for k, v in kw.items():
            setattr(m, k, v)
            okw[k] = v
        want = orc.cpn_post(*raw, (h, w), **okw)
        got = m.post(scores, locfou, ref, (h, w))
        assert len(got['scores'][0]) == len(want['scores'][0]), kw
        for k in ('contours', 'boxes', 'scores', 'locations', 'fourier', 'contour_proposals'):
            assert np.abs(got[k][0].cpu().numpy() - want[k][0].numpy()).max() < 1e-4, (kw, k)
        m.order, m.samples, m.refinement_iterations, m.score_thresh, m.nms_thresh = order, samples, 4, .9, .2
    # nothing above threshold -> empty lists with the reference's shapes
    got = m.post(scores - 100., locfou, ref, (h, w))
    assert got['contours'][0].shape == (0, samples, 2) and got['boxes'][0].shape == (0, 4)
    assert got['fourier'][0].shape == (0, order, 4) and got['classes'][0].dtype == torch.long
    # everything above threshold (P = h*w): selection order is raster order
    got = m.post(scores + 100., locfou, ref, (h, w), nms=False)
    hh, ww = scores.shape[1:]
    assert len(got['scores'][0]) == hh * ww
    loc = got['locations'][0].cpu().numpy()
    want = (raw[1][0].permute(1, 2, 0).reshape(-1, 2).numpy() + np.stack(np.meshgrid(np.arange(ww), np.arange(hh)), -1)
            .reshape(-1, 2)) * (w / ww)
    assert np.abs(loc - want).max() < 1e-4


def test_score_bounds():
    z = load_npz(MODEL_FIXTURES[0])
    m, (n, h, w, order, samples) = _post_model(z)
    raw = [torch.from_numpy(z[k]) for k in ('raw_scores', 'raw_locations', 'raw_refinement', 'raw_fourier')]
    g = torch.Generator().manual_seed(3)
    hh, ww = raw[0].shape[2:]
    upper = (torch.rand(n, 1, hh, ww, generator=g) > 0.5).float()
    lower = (torch.rand(n, 1, hh, ww, generator=g) > 0.98).float()
    want = orc.cpn_post(*raw, (h, w), order=order, samples=samples, scores_lower_bound=lower, scores_upper_bound=upper)
    got = m.post(raw[0][:, 0].contiguous().cuda(), torch.cat((raw[1], raw[3]), 1).permute(0, 2, 3, 1).contiguous().cuda(),
                 raw[2].permute(0, 2, 3, 1).contiguous().cuda(), (h, w), scores_lower_bound=lower.cuda(),
                 scores_upper_bound=upper.cuda())
    assert len(got['scores'][0]) == len(want['scores'][0]) > 0
    assert np.abs(got['contours'][0].cpu().numpy() - want['contours'][0].numpy()).max() < 1e-4
    # bounds given at the INPUT resolution (the tiled driver's mask crops): resized by cpn_resize_bilinear like
    # _equal_size does with F.interpolate (cpn.py:109-123)
    blocks = (torch.rand(n, 1, h // 8, w // 8, generator=g) > 0.4).float()
    upper_full = blocks.repeat_interleave(8, 2).repeat_interleave(8, 3)
    smooth = torch.rand(n, 1, h, w, generator=g) * 0.5
    want = orc.cpn_post(*raw, (h, w), order=order, samples=samples, scores_upper_bound=upper_full,
                        scores_lower_bound=smooth)
    got = m.post(raw[0][:, 0].contiguous().cuda(), torch.cat((raw[1], raw[3]), 1).permute(0, 2, 3, 1).contiguous().cuda(),
                 raw[2].permute(0, 2, 3, 1).contiguous().cuda(), (h, w), scores_upper_bound=upper_full.cuda(),
                 scores_lower_bound=smooth.cuda())
    assert len(got['scores'][0]) == len(want['scores'][0]) > 0
    assert np.abs(got['scores'][0].cpu().numpy() - want['scores'][0].numpy()).max() < 1e-6
    lib = cd._lib.load()
    src = torch.rand(2, 37, 53, 1, generator=g)
    dst = torch.empty(2, 20, 91, 1, device='cuda')
    cd._lib.check(lib.cpn_resize_bilinear(cd._lib.ptr(src.cuda()), 2, 37, 53, 1, cd._lib.ptr(dst), 20, 91,
                                          cd._lib.stream_ptr()))
    ref = torch.nn.functional.interpolate(src.permute(0, 3, 1, 2), (20, 91), mode='bilinear', align_corners=False)
    assert (dst.cpu().permute(0, 3, 1, 2) - ref).abs().max() < 1e-6


def test_fouriers2contours_golden_and_edges():
    z = load_npz('fouriers2contours')
    for order in (1, 5, 16):
        for samples in (32, 64, 128):
            t = f'o{order}_s{samples}'
            con, samp = cd.ops.cpn.fouriers2contours(torch.from_numpy(z[t + '/fourier']).cuda(),
                                                     torch.from_numpy(z[t + '/locations']).cuda(), samples=samples)
            assert con.shape == z[t + '/contours'].shape and samp.shape == (samples,)
            assert np.abs(con.cpu().numpy() - z[t + '/contours']).max() < 1e-3   # px (target 1e-3, gate 0.5)
    con, _ = cd.ops.cpn.fouriers2contours(torch.from_numpy(z['explicit/fourier']).cuda(),
                                          torch.from_numpy(z['explicit/locations']).cuda(),
                                          sampling=torch.from_numpy(z['explicit/sampling']).cuda())
    assert np.abs(con.cpu().numpy() - z['explicit/contours']).max() < 1e-3
    # P = 0 and P = 1, odd sample counts, list inputs (ops/cpn.py:60-63), leading batch dims
    f0, l0 = torch.zeros(0, 5, 4).cuda(), torch.zeros(0, 2).cuda()
    assert cd.ops.cpn.fouriers2contours(f0, l0, samples=32)[0].shape == (0, 32, 2)
    g = torch.Generator().manual_seed(5)
    f, l = torch.randn(3, 4, 7, 4, generator=g), torch.randn(3, 4, 2, generator=g)
    for s in (1, 2, 33, 65):
        got = cd.ops.cpn.fouriers2contours(f.cuda(), l.cuda(), samples=s)[0].cpu()
        want = orc.fouriers2contours(f, l, samples=s)[0]
        assert got.shape == want.shape and (got - want).abs().max() < 1e-3, s
    lst = cd.ops.cpn.fouriers2contours([f[0].cuda(), f[1].cuda()], [l[0].cuda(), l[1].cuda()], samples=16)
    assert len(lst) == 2 and lst[0][0].shape == (4, 16, 2)


def test_fouriers2contours_full_size_properties():
    """Config C5 (1e6 x order 16 x 128 samples): closure, linearity and a sampled comparison with the oracle."""
    P, order, samples = 1_000_000, 16, 128
    g = torch.Generator(device='cuda').manual_seed(0)
    f = torch.randn(P, order, 4, device='cuda', generator=g)
    loc = torch.rand(P, 2, device='cuda', generator=g) * 512
    con, _ = cd.ops.cpn.fouriers2contours(f, loc, samples=samples)
    assert con.shape == (P, samples, 2)
    assert (con[:, 0] - con[:, -1]).abs().max() < 2e-3           # t=0 and t=1 coincide (closed contour)
    con2, _ = cd.ops.cpn.fouriers2contours(f * 0.5, loc * 0.5, samples=samples)
    assert (con2 - con * 0.5).abs().max() < 1e-3                  # linear in (fourier, location)
    rows = torch.arange(0, P, 997, device='cuda')
    want = orc.fouriers2contours(f[rows].cpu(), loc[rows].cpu(), samples=samples)[0]
    assert (con[rows].cpu() - want).abs().max() < 1e-3


def _rand_boxes(n, g, extent=100., size=30.):
    xy = torch.rand(n, 2, generator=g) * extent
    return torch.cat((xy, xy + torch.rand(n, 2, generator=g) * size), 1)


def test_nms_matches_torchvision_exactly():
    import torchvision  # noqa: F401
    g = torch.Generator().manual_seed(0)
    for n in (1, 2, 33, 300, 1000, 5000):
        boxes, scores = _rand_boxes(n, g, extent=60. * (n ** .5) / 10), torch.rand(n, generator=g)
        if n >= 33:
            scores[3:12] = scores[3]                       # ties -> stable order
            boxes[5] = torch.tensor([5., 5., 5., 5.])      # zero-area pair: NaN IoU never suppresses
            boxes[6] = torch.tensor([5., 5., 5., 5.])
        for thr in (0.2, 0.5):
            want = torch.ops.torchvision.nms(boxes, scores, thr)
            got = cd.ops.cpn.nms(boxes.cuda(), scores.cuda(), thr).cpu()
            assert torch.equal(want, got), (n, thr)
    assert cd.ops.cpn.nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), .2).shape == (0,)
    b = torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 10.]])
    assert cd.ops.cpn.nms(b.cuda(), torch.tensor([.5, .5]).cuda(), 1.0).tolist() == [0, 1]   # IoU == thr keeps both


def test_batched_box_nmsi_segments_and_chunk_rule():
    g = torch.Generator().manual_seed(7)
    sizes = [0, 1, 700, 3000, 129]
    boxes = [_rand_boxes(n, g, extent=200.) for n in sizes]
    scores = [torch.rand(n, generator=g) for n in sizes]
    for bs in (None, 1024):
        want = orc.batched_box_nmsi([b.numpy() for b in boxes], [s.numpy() for s in scores], .3, batch_size=bs)
        got = cd.ops.cpn.batched_box_nmsi([b.cuda() for b in boxes], [s.cuda() for s in scores], .3, batch_size=bs)
        for w_, g_ in zip(want, got):
            assert np.array_equal(w_, g_.cpu().numpy()), bs


def test_remove_border_contours():
    g = torch.Generator().manual_seed(2)
    con = torch.rand(500, 32, 2, generator=g) * 80 - 8
    for sides in ((True, True, True, True), (False, True, False, True), (True, False, True, False)):
        t, r, b, l_ = sides
        off = torch.tensor([3., -2.])
        want = orc.remove_border_contours(con, (64, 64), 4, top=t, right=r, bottom=b, left=l_, offsets=off)
        got = cd.ops.cpn.remove_border_contours(con.cuda(), (64, 64), 4, top=t, right=r, bottom=b, left=l_,
                                                offsets=off)
        assert torch.equal(want, got.cpu())


def test_grid_nms_matches_torchvision_exactly():
    """Global stitch NMS (cpn_inference.py:405-408) at scale: the parallel grid algorithm returns exactly the greedy
    result (index for index) of torchvision's op."""
    import torchvision  # noqa: F401
    g = torch.Generator().manual_seed(3)
    for n, extent in ((1, 10.), (257, 100.), (5000, 400.), (60000, 3000.)):
        boxes, scores = _rand_boxes(n, g, extent=extent, size=40.), torch.rand(n, generator=g)
        if n > 100:
            scores[7:40] = scores[7]
            boxes[5] = boxes[6] = torch.tensor([5., 5., 5., 5.])
            boxes[11] = torch.tensor([0., 0., extent * 0.5, 30.])      # one long box
        for thr in (0.2, 0.5):
            want = torch.ops.torchvision.nms(boxes, scores, thr)
            got, rounds = cd.ops.cpn.nms_grid(boxes.cuda(), scores.cuda(), thr, return_rounds=True)
            assert torch.equal(want, got.cpu()), (n, thr)
            assert rounds >= 4
    # the public nms switches to the grid path for large inputs
    n = cd.ops.cpn.GRID_NMS_MIN + 10
    boxes, scores = _rand_boxes(n, g, extent=2000.), torch.rand(n, generator=g)
    assert torch.equal(torch.ops.torchvision.nms(boxes, scores, .2), cd.ops.cpn.nms(boxes.cuda(), scores.cuda(), .2).cpu())


def test_filter_contours_by_stitching_rule():
    g = torch.Generator().manual_seed(21)
    con = torch.rand(300, 1, 2, generator=g) * 70 + torch.rand(300, 16, 2, generator=g) * 6   # small blobs
    for ov in ([[8, 16], [8, 24]], [[0, 0], [0, 30]]):
        off = torch.tensor([-2., -3.])
        want = orc.filter_contours_by_stitching_rule(con, (64, 64), ov, offsets=off)
        got = cd.ops.cpn.filter_contours_by_stitching_rule(con.cuda(), (64, 64), torch.tensor(ov), offsets=off)
        assert torch.equal(want, got.cpu()) and 0 < int(want.sum()) < 300
        idx = cd.ops.cpn.filter_contours_by_stitching_rule(con.cuda(), (64, 64), torch.tensor(ov), offsets=off, indices=True)
        assert torch.equal(idx.cpu(), torch.where(want)[0])


def test_all_foreground_image_triggers_the_50000_chunk_rule():
    """Edge case of SURVEY appendix C.2: every pixel of a 256x256 head map is a proposal (P = 65 536 > NMS_BATCH_SIZE),
    so batched_box_nmsi runs NMS per 50 000-chunk and then over the survivors (ops/cpn.py:213-224)."""
    g = torch.Generator().manual_seed(5)
    n, h, w, order, samples = 1, 256, 256, 2, 8
    scores = torch.rand(n, 1, h, w, generator=g) * 4 + 3.0            # sigmoid > 0.95 everywhere
    locations = torch.randn(n, 2, h, w, generator=g)
    fourier = torch.randn(n, 4 * order, h, w, generator=g) * 2.0
    refinement = torch.tanh(torch.randn(n, 2, 2 * h, 2 * w, generator=g)) * 3
    want = orc.cpn_post(scores, locations, refinement, fourier, (2 * h, 2 * w), order=order, samples=samples)
    m = cd.models.CPN('CpnU22', order=order, samples=samples).cuda()
    locfou = torch.cat((locations, fourier), 1).permute(0, 2, 3, 1).contiguous().cuda()
    got = m.post(scores[:, 0].contiguous().cuda(), locfou, refinement.permute(0, 2, 3, 1).contiguous().cuda(),
                 (2 * h, 2 * w))
    assert len(m.post(scores[:, 0].contiguous().cuda(), locfou, refinement.permute(0, 2, 3, 1).contiguous().cuda(),
                      (2 * h, 2 * w), nms=False)['scores'][0]) == h * w
    assert len(got['scores'][0]) == len(want['scores'][0]) > 100
    assert np.abs(got['boxes'][0].cpu().numpy() - want['boxes'][0].numpy()).max() < 1e-3
    assert np.abs(got['scores'][0].cpu().numpy() - want['scores'][0].numpy()).max() < 1e-6
