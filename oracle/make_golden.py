"""TEST INFRASTRUCTURE ONLY -- mints the golden vectors under tests/golden from the *reference itself*.

Run in the build container (needs /root/reference):  ``python oracle/make_golden.py``

The reference has no golden vectors of its own for this path, so each fixture records what the
unmodified reference (imported through oracle/ref_shim.py, torch CPU fp32) computes for a seeded input and a seeded
synthetic state_dict.  State dicts are too large to commit (31-225 M parameters); fixtures store the seed of
``celldetection_b200.utils.synth_state_dict`` plus the six calibrated head tensors, which reproduces the exact
state_dict anywhere.  While minting, the oracle restatement (oracle/cpn_oracle.py) is checked against the reference
and the script aborts on any mismatch.
"""
import json
import os
import sys

os.environ.setdefault('TORCHDYNAMO_DISABLE', '1')   # ops/boxes.py wraps get_iou_voting in torch.compile; no OpenMP toolchain here
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
import cpn_oracle as orc  # noqa: E402
from celldetection_b200.utils.synth import synth_state_dict, calibrate_heads_  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
CALIB_KEYS = [f'core.{h}_head.block.4.{p}' for h in ('score', 'fourier', 'location', 'refinement')
              for p in ('weight', 'bias')]

FOURIER_STD, LOCATION_STD = 1.0, 0.5   # small objects so that NMS keeps a useful number of detections

# name, arch, N, H, W, seed, fg_fraction, ctor kwargs (variant), attributes set after construction
VARIANT_CASES = [
    ('cpnu22_c4_unc_b6', 'CpnU22', 2, 96, 128, 6, 0.2,
     dict(classes=4, uncertainty_head=True, refinement_buckets=6), dict(certainty_thresh=0.5, uncertainty_nms=True)),
    ('cpnresnext101unet_unc_b4', 'CpnResNeXt101UNet', 1, 128, 128, 7, 0.2,
     dict(uncertainty_head=True, refinement_buckets=4), dict(certainty_thresh=0.47)),
    ('cpnresnet18fpn_c3_b2', 'CpnResNet18FPN', 1, 128, 128, 8, 0.2,
     dict(classes=3, refinement_buckets=2), dict()),
    # shape-changing head options (models/cpn.py:177-234): per-head kernel sizes, narrowed ReadOut mid widths, strided heads
    ('cpnu22_k357_mid64', 'CpnU22', 1, 128, 128, 9, 0.2,
     dict(kernel_size_score=3, kernel_size_location=5, kernel_size_fourier=5, kernel_size_refinement=3,
          contour_head_channels=64), dict()),
    ('cpnresnet18fpn_mid_stride2', 'CpnResNet18FPN', 2, 128, 160, 10, 0.3,
     dict(contour_head_channels=128, refinement_head_channels=64, contour_head_stride=2, refinement_head_stride=2,
          kernel_size_refinement=5, backbone_kwargs=dict(fpn_channels=128)), dict()),
]


def variant_key(arch, ctor):
    return arch + ''.join(f'|{k}={v}' for k, v in sorted(ctor.items()))


MODEL_CASES = [
    # name, arch, N, H, W, seed, fg_fraction, model kwargs
    ('cpnu22_n1_128', 'CpnU22', 1, 128, 128, 1, 0.3, {}),
    ('cpnresnet18fpn_n2_128', 'CpnResNet18FPN', 2, 128, 128, 2, 0.2, {}),
    ('cpnresnext101unet_n1_128', 'CpnResNeXt101UNet', 1, 128, 128, 3, 0.2, {}),
    ('cpnu22_n2_96x160_s64', 'CpnU22', 2, 96, 160, 4, 0.2, dict(samples=64)),
]


def spec_of(ref_model):
    return OrderedDict((k, tuple(v.shape)) for k, v in ref_model.state_dict().items())


def build_state_dict(cd, arch, seed, fg_fraction, calib_x, ctor=None):
    model = getattr(cd.models, arch)(3, **(ctor or {})).eval()
    sd = synth_state_dict(spec_of(model), seed=seed)
    # torchvision's FeaturePyramidNetwork._load_from_state_dict renames 'inner_blocks.N.weight' for version < 2
    # state dicts; carry the module versions over so the reference loads our plain OrderedDict verbatim.
    sd._metadata = model.state_dict()._metadata

    def core_fn(x, sd_):
        model.load_state_dict(sd_)
        with torch.no_grad():
            scores, locations, refinement, fourier, unc = model.core(x)
        return dict(scores=scores, locations=locations, fourier=fourier, refinement=refinement, uncertainty=unc)

    calibrate_heads_(sd, core_fn, calib_x, fg_fraction=fg_fraction, fourier_std=FOURIER_STD, location_std=LOCATION_STD)
    if model.score_channels > 2:      # softmax scoring: centre every class logit, then favour the background class
        mu = core_fn(calib_x, sd)['scores'].mean((0, 2, 3))
        b = sd['core.score_head.block.4.bias'].clone() - mu
        b[0] += 2.0
        sd['core.score_head.block.4.bias'] = b
    model.load_state_dict(sd)
    return model, sd


def to_np(v):
    return v.detach().cpu().numpy()


def check_close(name, a, b, tol=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if a.size:
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)
        assert err <= tol, f'oracle != reference for {name}: rel err {err}'


def mint_model_case(cd, name, arch, n, h, w, seed, fg, mkw):
    torch.manual_seed(seed)
    x = torch.rand(n, 3, h, w)
    model, sd = build_state_dict(cd, arch, seed, fg, x[:1])
    for k, v in mkw.items():
        setattr(model, k, v)
    offsets = torch.tensor([[7., 3.]] * n) if n > 1 else None
    with torch.no_grad():
        scores, locations, refinement, fourier, _ = model.core(x)
        out = model(x) if offsets is None else model(x, offsets=offsets)
        out_nonms = model(x, nms=False)
    # ---- pin the oracle against the reference ----
    o_scores, o_loc, o_ref, o_fou = orc.cpn_core(x, sd, arch)
    for nm, a, b in (('scores', o_scores, scores), ('locations', o_loc, locations), ('refinement', o_ref, refinement),
                     ('fourier', o_fou, fourier)):
        check_close(f'{name}/{nm}', to_np(a), to_np(b), 1e-5)
    o_out = orc.cpn_forward(x, sd, arch, offsets=offsets, order=model.order, samples=model.samples)
    for k in ('contours', 'boxes', 'scores', 'locations', 'fourier', 'contour_proposals'):
        for i in range(n):
            check_close(f'{name}/out/{k}/{i}', to_np(o_out[k][i]), to_np(out[k][i]), 1e-5)
    arrays = dict(x=to_np(x), raw_scores=to_np(scores), raw_locations=to_np(locations), raw_refinement=to_np(refinement),
                  raw_fourier=to_np(fourier))
    if offsets is not None:
        arrays['offsets'] = to_np(offsets)
    for k in CALIB_KEYS:
        arrays['calib/' + k] = to_np(sd[k])
    for i in range(n):
        for k in ('contours', 'boxes', 'scores', 'classes', 'locations', 'fourier', 'contour_proposals'):
            arrays[f'out/{i}/{k}'] = to_np(out[k][i])
        arrays[f'nonms_count/{i}'] = np.array(len(out_nonms['scores'][i]))
    arrays['meta'] = np.array([n, h, w, seed, model.order, model.samples], dtype=np.int64)
    arrays['fg_fraction'] = np.array(fg)
    np.savez_compressed(os.path.join(GOLDEN, f'model_{name}.npz'), arch=np.array(arch), **arrays)
    print(f'{name}: proposals {[int(arrays[f"nonms_count/{i}"]) for i in range(n)]} kept '
          f'{[len(out["scores"][i]) for i in range(n)]}')


def mint_variant_case(cd, name, arch, n, h, w, seed, fg, ctor, attrs):
    """Variant models (classes > 2, uncertainty head, bucketed refinement): same recipe as ``mint_model_case``."""
    torch.manual_seed(seed)
    x = torch.rand(n, 3, h, w)
    model, sd = build_state_dict(cd, arch, seed, fg, x[:1], ctor)
    for k, v in attrs.items():
        setattr(model, k, v)
    offsets = torch.tensor([[7., 3.]] * n) if n > 1 else None
    with torch.no_grad():
        scores, locations, refinement, fourier, unc = model.core(x)
        out = model(x) if offsets is None else model(x, offsets=offsets)
        out_nonms = model(x, nms=False)
    core_kw = {k: ctor[k] for k in orc.CORE_KW if k in ctor}
    o = orc.cpn_core5(x, sd, arch, **core_kw)
    for nm, a, b in zip(('scores', 'locations', 'refinement', 'fourier', 'uncertainty'), o,
                        (scores, locations, refinement, fourier, unc)):
        if b is not None:
            check_close(f'{name}/{nm}', to_np(a), to_np(b), 1e-5)
    okw = dict(offsets=offsets, order=model.order, samples=model.samples, certainty_thresh=model.certainty_thresh,
               uncertainty_nms=model.uncertainty_nms, **core_kw)
    o_out = orc.cpn_forward(x, sd, arch, **okw)
    keys = ['contours', 'boxes', 'scores', 'classes', 'locations', 'fourier', 'contour_proposals']
    if unc is not None:
        keys.append('box_uncertainties')
    for k in keys:
        for i in range(n):
            assert len(o_out[k][i]) == len(out[k][i]), (name, k, i, len(o_out[k][i]), len(out[k][i]))
            check_close(f'{name}/out/{k}/{i}', to_np(o_out[k][i]), to_np(out[k][i]), 1e-5)
    arrays = dict(x=to_np(x), raw_scores=to_np(scores), raw_locations=to_np(locations), raw_refinement=to_np(refinement),
                  raw_fourier=to_np(fourier))
    if unc is not None:
        arrays['raw_uncertainty'] = to_np(unc)
    if offsets is not None:
        arrays['offsets'] = to_np(offsets)
    calib = list(CALIB_KEYS) + ([f'core.uncertainty_head.block.4.{p}' for p in ('weight', 'bias')] if unc is not None else [])
    for k in calib:
        arrays['calib/' + k] = to_np(sd[k])
    for i in range(n):
        for k in keys:
            arrays[f'out/{i}/{k}'] = to_np(out[k][i])
        arrays[f'nonms_count/{i}'] = np.array(len(out_nonms['scores'][i]))
    arrays['meta'] = np.array([n, h, w, seed, model.order, model.samples], dtype=np.int64)
    arrays['ctor'] = np.array(json.dumps(ctor))
    arrays['attrs'] = np.array(json.dumps(attrs))
    arrays['spec_key'] = np.array(variant_key(arch, ctor))
    np.savez_compressed(os.path.join(GOLDEN, f'model_{name}.npz'), arch=np.array(arch), **arrays)
    print(f'{name}: proposals {[int(arrays[f"nonms_count/{i}"]) for i in range(n)]} kept '
          f'{[len(out["scores"][i]) for i in range(n)]} classes {sorted(set(np.concatenate([to_np(c) for c in out["classes"]]).tolist()))}')


def mint_f2c(cd):
    arrays = {}
    g = torch.Generator().manual_seed(11)
    for order in (1, 5, 16):
        for samples in (32, 64, 128):
            P = 37
            f = torch.randn(P, order, 4, generator=g)
            loc = torch.rand(P, 2, generator=g) * 512
            con, _ = cd.ops.cpn.fouriers2contours(f, loc, samples=samples)
            o_con, _ = orc.fouriers2contours(f, loc, samples=samples)
            check_close(f'f2c/{order}/{samples}', to_np(o_con), to_np(con), 1e-6)
            tag = f'o{order}_s{samples}'
            arrays[tag + '/fourier'], arrays[tag + '/locations'], arrays[tag + '/contours'] = to_np(f), to_np(loc), to_np(con)
    # explicit per-proposal sampling (ops/cpn.py:67-71)
    f = torch.randn(9, 5, 4, generator=g)
    loc = torch.rand(9, 2, generator=g) * 64
    samp = torch.sort(torch.rand(9, 48, generator=g), -1).values
    con, _ = cd.ops.cpn.fouriers2contours(f, loc, sampling=samp)
    arrays['explicit/fourier'], arrays['explicit/locations'] = to_np(f), to_np(loc)
    arrays['explicit/sampling'], arrays['explicit/contours'] = to_np(samp), to_np(con)
    np.savez_compressed(os.path.join(GOLDEN, 'fouriers2contours.npz'), **arrays)
    print('fouriers2contours: ok')


def check_stitching_rule(cd):
    g = torch.Generator().manual_seed(21)
    con = torch.rand(300, 1, 2, generator=g) * 70 + torch.rand(300, 16, 2, generator=g) * 6   # small blobs
    for ov in ([[8, 16], [8, 24]], [[0, 0], [0, 30]]):
        want = cd.ops.filter_contours_by_stitching_rule(con, (64, 64), torch.tensor(ov), offsets=torch.tensor([-2., -3.]))
        got = orc.filter_contours_by_stitching_rule(con, (64, 64), ov, offsets=torch.tensor([-2., -3.]))
        assert torch.equal(want, got)
    print('stitching rule: ok')


def mint_tiling(cd):
    check_stitching_rule(cd)
    arrays = {}
    cases = [((640, 896), (256, 256), (192, 192)), ((512, 512), (512, 512), (384, 384)), ((100, 300), (128, 128), (96, 96)),
             ((2048, 2048), (512, 512), (384, 384)), ((1000, 777), (256, 192), (200, 100))]
    for i, (size, crop, strides) in enumerate(cases):
        sl, ov, shape = cd.get_tiling_slices(size, crop, strides, return_overlaps=True)
        sl, ov = list(sl), list(ov)
        arrays[f'{i}/args'] = np.array([size, crop, strides])
        arrays[f'{i}/slices'] = np.array([[[s.start, s.stop] for s in t] for t in sl])
        arrays[f'{i}/overlaps'] = np.array(ov).reshape(len(sl), len(size), 2)
        arrays[f'{i}/shape'] = np.array(shape)
        o_sl, o_ov, o_shape = orc.get_tiling_slices(size, crop, strides)
        assert [[(s.start, s.stop) for s in t] for t in o_sl] == [[(int(s.start), int(s.stop)) for s in t] for t in sl]
        assert np.array_equal(o_ov, arrays[f'{i}/overlaps']) and tuple(o_shape) == tuple(shape)
    np.savez_compressed(os.path.join(GOLDEN, 'tiling.npz'), **arrays)
    print('tiling: ok')


def mint_apply_model(cd):
    """Reference ``apply_model`` (FakeTrainer) on a small synthetic uint8 image, CpnU22, crop 64 / stride 48."""
    import importlib
    cs = importlib.import_module('celldetection_scripts.cpn_inference')
    cs = sys.modules['celldetection_scripts.cpn_inference']
    seed = 5
    torch.manual_seed(seed)
    calib = torch.rand(1, 3, 64, 64)
    model, sd = build_state_dict(cd, 'CpnU22', seed, 0.05, calib)
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, size=(150, 200, 3)).astype(np.uint8)
    lit = cd.models.LitCpn(model)
    lit.eval()
    lit.max_imsize = None
    res = cs.apply_model(img, [lit], ref_shim.FakeTrainer(), crop_size=(64, 64), strides=(48, 48),
                         model_kwargs_list=[{}], batch_size=2, verbose=False)
    o_res = orc.apply_model(img, sd, 'CpnU22', 64, 48, border_removal=4)
    assert len(o_res['scores']) == len(res['scores']), (len(o_res['scores']), len(res['scores']))
    for k in ('contours', 'boxes', 'scores', 'locations'):
        check_close(f'apply_model/{k}', to_np(o_res[k]), to_np(res[k]), 1e-5)
    arrays = dict(img=img, meta=np.array([seed, 64, 48, 4]))
    for k in CALIB_KEYS:
        arrays['calib/' + k] = to_np(sd[k])
    for k in ('contours', 'boxes', 'scores', 'classes', 'locations', 'fourier', 'contour_proposals'):
        arrays['out/' + k] = to_np(res[k])
    # test-time repetitions: every tile 3x, repetition r through a host transform of the crop (TileLoader :85-91,114-118)
    res_t = cs.apply_model(img, [lit], ref_shim.FakeTrainer(), crop_size=(64, 64), strides=(48, 48), reps=3,
                           transforms=orc.tta_example_transform, model_kwargs_list=[{}], batch_size=2, verbose=False)
    o_t = orc.apply_model(img, sd, 'CpnU22', 64, 48, border_removal=4, reps=3, transforms=orc.tta_example_transform)
    assert len(o_t['scores']) == len(res_t['scores']) != len(res['scores']), (len(o_t['scores']), len(res_t['scores']))
    for k in ('contours', 'boxes', 'scores', 'locations'):
        check_close(f'apply_model/tta/{k}', to_np(o_t[k]), to_np(res_t[k]), 1e-5)
    for k in ('contours', 'boxes', 'scores', 'classes', 'locations', 'fourier', 'contour_proposals'):
        arrays['tta/' + k] = to_np(res_t[k])
    np.savez_compressed(os.path.join(GOLDEN, 'apply_model_cpnu22.npz'), **arrays)
    print(f'apply_model: {len(res["scores"])} stitched detections, {len(res_t["scores"])} with 3 repetitions')


def mint_apply_ensemble(cd):
    """Reference ``apply_model`` with TWO CpnU22 models, ``min_vote=2`` and a ``mask`` (upper score bound; tiles with an
    empty mask crop are skipped): pins the oracle's ensemble / voting / mask path (cpn_inference.py:93-111, 417-427)."""
    import importlib
    importlib.import_module('celldetection_scripts.cpn_inference')
    cs = sys.modules['celldetection_scripts.cpn_inference']
    seeds = (5, 5)     # the second member is the first with perturbed heads: similar, not identical, detections
    lits, sds = [], []
    for j, seed in enumerate(seeds):
        torch.manual_seed(seed)
        calib = torch.rand(1, 3, 64, 64)
        model, sd = build_state_dict(cd, 'CpnU22', seed, 0.08, calib)
        if j == 1:
            sd['core.location_head.block.4.weight'] = sd['core.location_head.block.4.weight'] * 1.3
            sd['core.fourier_head.block.4.weight'] = sd['core.fourier_head.block.4.weight'] * 0.85
            sd['core.score_head.block.4.bias'] = sd['core.score_head.block.4.bias'] - 0.4
            model.load_state_dict(sd)
        lit = cd.models.LitCpn(model)
        lit.eval()
        lit.max_imsize = None
        lits.append(lit)
        sds.append(sd)
    rng = np.random.RandomState(17)
    img = rng.randint(0, 256, size=(150, 200, 3)).astype(np.uint8)
    mask = np.zeros((150, 200), dtype=bool)
    mask[10:140, 30:120] = True              # leaves the right-most tile column empty
    mask[40:60, 50:70] = False
    arrays = dict(img=img, mask=mask, meta=np.array([64, 48, 4]), seeds=np.array(seeds))
    # NOTE: with min_vote > 1 the reference raises IndexError as soon as the vote removes a box (it stores the filtered
    # votes in the result dict BEFORE applying the keep indices to it, cpn_inference.py:421-423), so the driver-level
    # fixture uses min_vote=1 and the vote itself is pinned at op level below.
    g = torch.Generator().manual_seed(23)
    ctr = torch.rand(60, 2, generator=g) * 300
    bx = torch.cat([torch.cat((c - 8 + torch.randn(k, 2, generator=g) * 2, c + 8 + torch.randn(k, 2, generator=g) * 2), 1)
                    for c, k in zip(ctr, torch.randint(1, 5, (60,), generator=g).tolist())], 0)
    bx[3, 2:] = bx[3, :2]          # a zero-area box: NaN vote, never kept
    keep, votes = cd.ops.filter_by_box_voting(bx, 0.2, 2, return_votes=True)
    o_keep, o_votes = orc.filter_by_box_voting(bx, 0.2, 2)
    assert torch.equal(keep.long(), o_keep.long()) and 0 < len(keep) < len(bx)
    check_close('votes', to_np(o_votes), to_np(votes), 1e-6)
    arrays.update({'voting/boxes': to_np(bx), 'voting/keep': to_np(keep), 'voting/votes': to_np(votes)})
    for tag, kw in (('vote1', dict(min_vote=1)),):
        res = cs.apply_model(img, lits, ref_shim.FakeTrainer(), mask=mask, crop_size=(64, 64), strides=(48, 48),
                             model_kwargs_list=[{}, {}], batch_size=1, verbose=False, **kw)
        o_res = orc.apply_models(img, sds, ['CpnU22'] * 2, 64, 48, border_removal=4, mask=mask, **kw)
        assert len(o_res['scores']) == len(res['scores']) > 0, (tag, len(o_res['scores']), len(res['scores']))
        keys = ['contours', 'boxes', 'scores', 'locations'] + (['votes'] if 'votes' in res else [])
        for k in keys:
            check_close(f'apply_ensemble/{tag}/{k}', to_np(o_res[k]), to_np(res[k]), 1e-5)
        for k in ['contours', 'boxes', 'scores', 'classes', 'locations', 'fourier', 'contour_proposals'] + keys[4:]:
            arrays[f'{tag}/{k}'] = to_np(res[k])
        print(f'apply_model ensemble {tag}: {len(res["scores"])} detections')
    for j, sd in enumerate(sds):
        for k in CALIB_KEYS:
            arrays[f'calib{j}/' + k] = to_np(sd[k])
    np.savez_compressed(os.path.join(GOLDEN, 'apply_model_ensemble.npz'), **arrays)


def synth_contours(rng, K, S, H, W, radius=(2., 14.)):
    """CPN-like closed contours (random low-order Fourier shapes, inclusive linspace sampling like ops/cpn.py:67-81)."""
    t = np.linspace(0, 1, S)
    out = np.zeros((K, S, 2), np.float32)
    for k in range(K):
        order, r = rng.randint(1, 6), rng.uniform(*radius)
        xy = np.array([rng.rand() * W, rng.rand() * H])[None] + sum(
            rng.randn(2)[None] * r / (j + 1) * np.cos(2 * np.pi * (j + 1) * t)[:, None] +
            rng.randn(2)[None] * r / (j + 1) * np.sin(2 * np.pi * (j + 1) * t)[:, None] for j in range(order))
        out[k] = xy
    return out


def mint_contours2labels(cd):
    """cd.data.contours2labels of the reference (cv2.drawContours fill + greedy channel assignment) on seeded contour
    sets: sparse, dense (many channels), image-border clipping, degenerate (collapsed) contours, odd sample counts."""
    import c2l_oracle as c2l
    rng = np.random.RandomState(31)
    arrays = {}
    cases = [('sparse', 60, 32, 160, 200, (2., 9.)), ('dense', 150, 16, 64, 96, (3., 12.)),
             ('border', 80, 64, 90, 120, (6., 25.)), ('odd', 40, 21, 70, 70, (1., 6.))]
    for name, K, S, H, W, rad in cases:
        con = synth_contours(rng, K, S, H, W, rad)
        if name == 'odd':
            con[3] = con[3, :1]                         # collapsed to a point
            con[7, :, 1] = con[7, 0, 1]                 # horizontal segment
            con[11] = np.round(con[11]) + 0.5           # exact .5 coordinates: np.round is half-to-even
        want = cd.data.contours2labels(con.copy(), (H, W))
        got = c2l.contours2labels(con.copy(), (H, W))
        assert want.shape == got.shape and np.array_equal(want, got), name
        arrays[f'{name}/contours'] = con
        arrays[f'{name}/size'] = np.array([H, W])
        arrays[f'{name}/labels'] = want.astype(np.int16)
        print(f'contours2labels {name}: {K} contours -> {want.shape[2]} channels')
    # unrounded / gap variants
    con = synth_contours(rng, 50, 32, 80, 80, (2., 10.))
    for tag, kw in (('noround', dict(rounded=False)), ('gap0', dict(gap=0)), ('depth3', dict(initial_depth=3))):
        want = cd.data.contours2labels(con.copy(), (80, 80), **kw)
        got = c2l.contours2labels(con.copy(), (80, 80), **kw)
        assert want.shape == got.shape and np.array_equal(want, got), tag
        arrays[f'{tag}/labels'] = want.astype(np.int16)
    arrays['variants/contours'] = con
    # resolve_label_channels (data/cpn.py:361-398) on the label images above: overlap pixels filled by dilation sweeps
    for name in ('sparse', 'dense', 'border', 'odd'):
        lab = arrays[f'{name}/labels'].astype(np.int32)
        want = cd.data.resolve_label_channels(lab)
        got = c2l.resolve_label_channels(lab)
        assert want.shape == got.shape and np.array_equal(want, got), name
        arrays[f'{name}/flat'] = want.astype(np.int16)
    one = arrays['sparse/labels'].astype(np.int32)[..., :1]                   # no overlaps at all: plain max
    assert np.array_equal(cd.data.resolve_label_channels(one), c2l.resolve_label_channels(one))
    np.savez_compressed(os.path.join(GOLDEN, 'contours2labels.npz'), **arrays)


def mint_preprocess(cd):
    """Slide preprocessing (cpn_inference.py:196-222): the reference's own ``cd.data.normalize_percentile`` (float result:
    its uint8 conversion is skimage's, absent here) pins the oracle's percentile / clip / scale step; the oracle's full
    chain (uint8 conversion, gamma, contrast restated from the third-party sources) is stored for the GPU test."""
    import preprocess_oracle as po
    rng = np.random.RandomState(3)
    arrays = {}
    img16 = rng.gamma(2.0, 900, size=(157, 211)).clip(0, 65535).astype(np.uint16)
    img8 = rng.gamma(2.0, 30, size=(96, 130, 3)).clip(0, 255).astype(np.uint8)
    for name, img in (('u16', img16), ('u8', img8)):
        arrays[f'{name}/img'] = img
        for j, pct in enumerate((99.9, 98.0, (0.5, 99.0))):
            ref = cd.data.normalize_percentile(img, pct, to_uint8=False)
            got = po.normalize_percentile(img, pct, to_uint8=False)
            assert np.array_equal(np.asarray(ref), got), (name, pct)
            arrays[f'{name}/pct{j}'] = np.asarray(pct, dtype=np.float64).reshape(-1)
            arrays[f'{name}/norm{j}'] = po.normalize_percentile(img, pct)
        chains = ((0.8, 1., 0., None, 0), (1., 1.3, 0.1, 99.5, 0), (1.4, 0.7, -0.05, None, 0), (1., 1., 0., None, 1),
                  (0.9, 1.2, 0.05, 99., 1))
        for j, (gamma, con, bri, pct, gray) in enumerate(chains):
            arrays[f'{name}/chain{j}/params'] = np.array([gamma, con, bri, -1. if pct is None else pct, gray])
            arrays[f'{name}/chain{j}/out'] = po.preprocess(img, gamma, con, bri, pct, grayscale=bool(gray))
    rgba = rng.randint(0, 256, size=(64, 80, 4)).astype(np.uint8)
    import cv2
    assert np.array_equal(po.rgb2gray(rgba), cv2.cvtColor(rgba, cv2.COLOR_RGBA2GRAY))
    assert np.array_equal(po.rgb2gray(img8), cv2.cvtColor(img8, cv2.COLOR_RGB2GRAY))
    arrays['rgba/img'] = rgba
    arrays['rgba/chain0/params'] = np.array([1.2, 1., 0., -1., 1])
    arrays['rgba/chain0/out'] = po.preprocess(rgba, 1.2, 1., 0., None, grayscale=True)
    np.savez_compressed(os.path.join(GOLDEN, 'preprocess.npz'), **arrays)
    print('preprocess: ok', sorted(arrays)[:4], '...')


def mint_keys(cd):
    """state_dict key -> shape of the three reference models (drop-in contract)."""
    out = {}
    for arch in ('CpnU22', 'CpnResNet18FPN', 'CpnResNeXt101UNet'):
        m = getattr(cd.models, arch)(3)
        out[arch] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    from celldetection_b200.models.graph import ARCHS as ALL_ARCHS
    for arch in ALL_ARCHS[3:]:           # the rest of the ResNet family: keys + a numerical pin of the oracle, no vectors
        m = getattr(cd.models, arch)(3).eval()
        out[arch] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        sd = synth_state_dict(spec_of(m), seed=3)
        sd._metadata = m.state_dict()._metadata
        m.load_state_dict(sd)
        torch.manual_seed(1)
        x = torch.rand(1, 3, 64, 64)
        with torch.no_grad():
            ref = m.core(x)
            got = orc.cpn_core(x, sd, arch)
        for nm, a, b in zip(('scores', 'locations', 'refinement', 'fourier'), got, ref):
            check_close(f'{arch}/{nm}', to_np(a), to_np(b), 1e-5)
        print(f'{arch}: oracle == reference on a 64x64 tile')
    for _, arch, _, _, _, _, _, ctor, _ in VARIANT_CASES:
        m = getattr(cd.models, arch)(3, **ctor)
        out[variant_key(arch, ctor)] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    with open(os.path.join(GOLDEN, 'state_dict_keys.json'), 'w') as f:
        json.dump(out, f)
    print('keys: ok', {k: len(v) for k, v in out.items()})


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    cd = ref_shim.import_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    which = sys.argv[1:] or ['keys', 'f2c', 'tiling', 'models', 'variants', 'apply', 'ensemble', 'labels', 'preprocess']
    if 'keys' in which:
        mint_keys(cd)
    if 'f2c' in which:
        mint_f2c(cd)
    if 'tiling' in which:
        mint_tiling(cd)
    if 'models' in which:
        for case in MODEL_CASES:
            mint_model_case(cd, *case)
    if 'variants' in which:
        for case in VARIANT_CASES:
            mint_variant_case(cd, *case)
    if 'apply' in which:
        mint_apply_model(cd)
    if 'ensemble' in which:
        mint_apply_ensemble(cd)
    if 'labels' in which:
        mint_contours2labels(cd)
    if 'preprocess' in which:
        mint_preprocess(cd)


if __name__ == '__main__':
    main()
