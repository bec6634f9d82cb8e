"""CPU: host-side logic of the product package -- state_dict contract, C-ABI surface, graph lowering, weight packing,
arena assignment, tiling, and the world_size-2 exchange step over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import celldetection_b200 as cd
from celldetection_b200 import _lib
from celldetection_b200.models import graph as G
from celldetection_b200.models import plan as PL
from conftest import ROOT
from helpers import key_spec, load_npz, fixture_ctor, VARIANT_FIXTURES


@pytest.mark.parametrize('arch', G.ARCHS)
def test_state_dict_keys_match_reference(arch):
    """Drop-in contract: same keys, same order, same shapes as the reference's state_dict."""
    model = getattr(cd.models, arch)(3)
    want = key_spec(arch)
    got = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    assert [k for k, _ in got] == list(want.keys())
    assert got == list(want.items())


@pytest.mark.parametrize('name', VARIANT_FIXTURES)
def test_variant_state_dict_keys_match_reference(name):
    """classes > 2 / uncertainty_head / refinement_buckets change the head widths and add ``core.uncertainty_head``
    between the fourier and refinement heads (models/cpn.py:177-234): keys, order and shapes as the reference."""
    z = load_npz(name)
    ctor, attrs = fixture_ctor(z)
    model = getattr(cd.models, str(z['arch']))(3, **ctor)
    want = key_spec(str(z['spec_key']))
    got = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    assert got == list(want.items())
    for k, v in attrs.items():
        assert hasattr(model, k)


def test_bucket_table_matches_reference_formula():
    """ops/cpn.py:238-255 on the default sampling: indices in [0, B), weights >= 0 and summing to 1 per sample."""
    from celldetection_b200.ops.cpn import bucket_table
    for samples, buckets in ((32, 2), (32, 6), (64, 4), (128, 12)):
        idx, w = bucket_table(samples, buckets, 'cpu')
        assert idx.shape == w.shape == (samples, 3) and idx.dtype == torch.int32
        assert int(idx.min()) >= 0 and int(idx.max()) < buckets
        assert float(w.min()) >= 0 and torch.allclose(w.sum(1), torch.ones(samples), atol=1e-6)
        t = torch.linspace(0, 1.0, samples) * buckets
        assert torch.equal(idx[:, 1].long(), t.long() % buckets)


def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, 'include', 'cpn_b200.h')) as f:
        header = f.read()
    declared = set(re.findall(r'\b(cpn_[a-z0-9_]+)\s*\(', header))
    declared -= {'cpn_plan_t', 'cpn_op_t', 'cpn_view_t'}
    assert declared == set(_lib.SYMBOLS.keys()), declared ^ set(_lib.SYMBOLS.keys())
    lib = _lib.load()                      # raises if the .so is missing or a symbol is not exported
    assert lib.cpn_abi_version() == _lib.ABI_VERSION
    assert lib.cpn_last_error() is not None
    assert lib.cpn_select_workspace_bytes(1 << 20) > 0


def test_struct_layout_matches_header():
    import ctypes
    assert ctypes.sizeof(_lib.View) == 40
    assert ctypes.sizeof(_lib.Op) == 8 + 3 * 40 + 16 + 4 * 4 + 8 * 4 + 8


def test_no_cpu_fallback():
    model = cd.models.CpnU22(3)
    with pytest.raises(RuntimeError):
        model(torch.rand(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        cd.ops.cpn.fouriers2contours(torch.zeros(3, 5, 4), torch.zeros(3, 2))
    with pytest.raises(NotImplementedError):
        cd.models.CpnU22(3, backbone_kwargs=dict(depth=4))
    with pytest.raises(NotImplementedError):
        cd.models.CpnU22(3).train()(torch.rand(1, 3, 64, 64), targets={})


def test_graph_flop_census():
    """Executed conv FLOPs per 3x512x512 (256x256 for U22) tile vs SURVEY 8a/8d: reference totals 201.67 / 2124.85 /
    2392.83 GFLOP minus the stated savings (1x1-upsample commute, dead FPN output convs)."""
    for arch, hw, ref_total, saved in (('CpnU22', 256, 201.67, 3.2), ('CpnResNet18FPN', 512, 2124.85, 6.3),
                                       ('CpnResNeXt101UNet', 512, 2392.83, 45.1)):
        g = G.trace(arch, 1, hw, hw)
        gf = G.conv_flops(g) / 1e9
        assert abs(gf - (ref_total - saved)) / ref_total < 0.004, (arch, gf)


def test_grouped_expansion_and_bn_folding_equal_conv():
    torch.manual_seed(0)
    for cg, c in ((8, 128), (16, 128), (64, 128)):
        groups = c // cg
        x = torch.randn(1, c, 9, 9)
        sd = {'c.weight': torch.randn(c, cg, 3, 3) * 0.1, 'bn.weight': torch.rand(c) + 0.5, 'bn.bias': torch.randn(c),
              'bn.running_mean': torch.randn(c) * 0.1, 'bn.running_var': torch.rand(c) + 0.5}
        ref = F.batch_norm(F.conv2d(x, sd['c.weight'], None, padding=1, groups=groups), sd['bn.running_mean'],
                           sd['bn.running_var'], sd['bn.weight'], sd['bn.bias'], False, 0., 1e-5)
        w, b = PL.fold_conv(sd, G.ConvParams(['c.weight'], [None], ['bn'], groups))
        wexp = PL.expand_grouped(w, groups)             # [cout, kslab, 3, 3]
        kslab, mode = PL.slab_of(c, c, groups)
        assert mode == 1 and wexp.shape[1] == kslab
        out = torch.zeros_like(ref)
        for n0 in range(0, c, 64):                      # what the kernels do per 64-wide output slab
            base = (n0 // kslab) * kslab
            out[:, n0:n0 + 64] = F.conv2d(x[:, base:base + kslab], wexp[n0:n0 + 64], b[n0:n0 + 64], padding=1)
        assert torch.allclose(out, ref, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize('arch', G.ARCHS)
def test_arena_assignment_has_no_live_overlap(arch):
    g = G.trace(arch, 2, 128, 128)
    offsets, total = PL.assign_arena(g, 2)
    roots = [(t, offsets[t.id]) for t in g.tensors if t.id in offsets]
    sizes = {t.id: PL._align(2 * t.h * t.w * (t.c if t.c % 4 == 0 else PL._align(t.c, 4)) * 2) for t, _ in roots}
    for i, (a, oa) in enumerate(roots):
        assert oa % 256 == 0 and oa + sizes[a.id] <= total
        for b, ob in roots[i + 1:]:
            live = not (a.last < b.first or b.last < a.first)
            overlap = not (oa + sizes[a.id] <= ob or ob + sizes[b.id] <= oa)
            assert not (live and overlap), (a.id, b.id)
    # every op reads tensors that were written before
    written = set()
    for op in g.ops:
        for t in (op.src, op.res):
            if t is not None:
                assert t.root()[0].id in written or t.id in written, op.name
        written.add(op.dst.id)
        written.add(op.dst.root()[0].id)


def test_get_tiling_slices_matches_reference_golden():
    z = load_npz('tiling')
    i = 0
    while f'{i}/args' in z.files:
        size, crop, strides = [tuple(int(v) for v in r) for r in z[f'{i}/args']]
        sl, ov, shape = cd.get_tiling_slices(size, crop, strides, return_overlaps=True)
        sl = list(sl)
        assert np.array_equal(np.array([[[s.start, s.stop] for s in t] for t in sl]), z[f'{i}/slices'])
        assert np.array_equal(np.array(list(ov)).reshape(len(sl), 2, 2), z[f'{i}/overlaps'])
        assert tuple(shape) == tuple(z[f'{i}/shape'])
        i += 1


def test_model_file_round_trip(tmp_path):
    m = cd.models.CpnU22(3, order=4, samples=48)
    f = str(tmp_path / 'm.pt')
    cd.save_fetchable_model(m, f)
    m2 = cd.load_model(f, map_location='cpu')
    assert type(m2).__name__ == 'CpnU22' and m2.order == 4 and m2.samples == 48
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from collections import OrderedDict
from celldetection_b200.inference import allgather_detections, canonical_order
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:' + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
KS = [int(v) for v in sys.argv[5].split(',')]
K = KS[rank]
g = torch.Generator().manual_seed(100 + rank)
res = OrderedDict(contours=torch.rand(K, 8, 2, generator=g), boxes=torch.rand(K, 4, generator=g),
                  scores=torch.rand(K, generator=g), classes=torch.ones(K, dtype=torch.long),
                  locations=torch.rand(K, 2, generator=g), fourier=torch.rand(K, 5, 4, generator=g),
                  contour_proposals=torch.rand(K, 8, 2, generator=g),
                  box_uncertainties=torch.rand(K, 4, generator=g),      # models with an uncertainty head carry this key
                  order_key=torch.stack((torch.arange(K).float() * 2 + rank, torch.arange(K).float()), 1))
out = allgather_detections(res)
T = sum(KS)
assert out['scores'].shape[0] == T and out['classes'].dtype == torch.long
assert list(out.keys()) == list(res.keys()) and out['box_uncertainties'].shape == (T, 4)
assert out['contours'].shape == (T, 8, 2) and out['fourier'].shape == (T, 5, 4)
for r, (a, b) in enumerate(((0, KS[0]), (KS[0], T))):
    gg = torch.Generator().manual_seed(100 + r)
    kk = b - a
    want = torch.rand(kk, 8, 2, generator=gg)
    assert torch.equal(out['contours'][a:b], want), r
can = canonical_order(out)                       # tiles were dealt round-robin: rank 0 -> 0,2,4  rank 1 -> 1,3,5,7,9
if KS == [3, 5]:
    assert can['order_key'][:, 0].tolist() == [0., 1., 2., 3., 4., 5., 7., 9.]
    assert torch.equal(can['contours'][1], out['contours'][3])
assert can['scores'].shape[0] == T
torch.save({k: v for k, v in out.items()}, sys.argv[4] + f'.{rank}')
dist.destroy_process_group()
'''


@pytest.mark.parametrize('counts', ['3,5', '0,5', '4,0', '0,0'])
def test_allgather_detections_world2_gloo(tmp_path, counts):
    """N > 1 exchange step on CPU: both ranks end up with the identical rank-ordered concatenation --
    including when one rank or every rank has no detections (fewer tiles than ranks, blank tiles, masked-out tiles)."""
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    port = str(29500 + (os.getpid() * 7 + sum(int(c) for c in counts.split(',')) * 13 + len(counts)) % 2000)
    out = str(tmp_path / 'res')
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), out, counts]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    a, b = torch.load(out + '.0'), torch.load(out + '.1')
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_im2col_stem_trace_keeps_flops_and_shapes():
    """The tensor-core engine rewrites the C_in = 3 stem as prep(im2col) + 1x1 conv: same arithmetic, same FLOP census."""
    for arch, hw in (('CpnU22', 64), ('CpnResNeXt101UNet', 128)):
        a, b = G.trace(arch, 2, hw, hw), G.trace(arch, 2, hw, hw, stem_im2col=True)
        assert G.conv_flops(a) == G.conv_flops(b)
        assert len(a.ops) == len(b.ops)
        prep, stem = b.ops[0], b.ops[1]
        k, cin = prep.im2col
        assert prep.kind == 'prep' and stem.k == 1 and stem.im2col == (k, cin)
        assert stem.src.c % 64 == 0 and stem.src.c >= k * k * cin
        assert (stem.dst.h, stem.dst.w) == (a.ops[1].dst.h, a.ops[1].dst.w)
        assert list(a.spec.keys()) == list(b.spec.keys())


def test_synth_state_dict_is_deterministic_and_complete():
    from celldetection_b200.utils.synth import synth_state_dict
    spec = key_spec('CpnResNet18FPN')
    a, b = synth_state_dict(spec, seed=3), synth_state_dict(spec, seed=3)
    assert list(a.keys()) == list(spec.keys())
    assert all(torch.equal(a[k], b[k]) and tuple(a[k].shape) == tuple(spec[k]) for k in spec)
    c = synth_state_dict(spec, seed=4)
    assert not torch.equal(a['core.score_head.block.0.weight'], c['core.score_head.block.0.weight'])


def test_tile_bounds_follow_the_reference_tile_loader():
    """TileLoader.__getitem__ (cpn_inference.py:93-111): mask crop -> upper bound, clipped point-mask crop -> lower bound
    (and upper, when exclusive); a tile whose crop is empty is skipped."""
    from celldetection_b200.inference import _tile_bounds
    mask = np.zeros((8, 8), bool)
    mask[:4, :4] = True
    pm = np.zeros((8, 8), np.float32)
    pm[1, 1] = 3.
    sl_in, sl_out = (slice(0, 4), slice(0, 4)), (slice(4, 8), slice(4, 8))
    up, lo = _tile_bounds(mask, None, False, sl_in)
    assert lo is None and up.shape == (4, 4, 1) and up.dtype == np.float32 and up.all()
    assert _tile_bounds(mask, None, False, sl_out) is False
    up, lo = _tile_bounds(None, pm, False, sl_in)
    assert up is None and lo.shape == (4, 4, 1) and lo.max() == 1. and lo.sum() == 1.
    up, lo = _tile_bounds(mask, pm, True, sl_in)
    assert up is lo
    assert _tile_bounds(mask, pm, False, sl_out) is False
    assert _tile_bounds(None, None, False, sl_in) == (None, None)


def test_trace_variants_widen_the_heads_like_the_reference():
    """models/cpn.py:177-234: score head -> `classes` channels (classes > 2), refinement head -> 2 * buckets channels,
    uncertainty head = a fourth ReadOut with 4 sigmoid outputs on the head features; the merged 7x7 head convolution
    grows to 4 x C_head and every projection reads its own C_head-wide slice."""
    g0 = G.trace('CpnResNeXt101UNet', 1, 128, 128)
    g1 = G.trace('CpnResNeXt101UNet', 1, 128, 128, score_channels=5, refinement_buckets=6, uncertainty_head=True)
    assert list(g0.outputs) == ['scores', 'locfou', 'refinement']
    assert list(g1.outputs) == ['scores', 'locfou', 'refinement', 'uncertainty']
    assert (g1.outputs['scores'].c, g1.outputs['refinement'].c, g1.outputs['uncertainty'].c) == (5, 12, 4)
    heads0 = [o for o in g0.ops if o.name == 'heads.block.0'][0]
    heads1 = [o for o in g1.ops if o.name == 'heads.block.0'][0]
    assert heads0.dst.c == 3 * 256 and heads1.dst.c == 4 * 256
    projs = [o for o in g1.ops if o.kind == 'proj' and o.src is heads1.dst]
    assert [(o.cin_off, o.cin, o.dst.c, o.act) for o in projs] == [(0, 256, 5, 'none'), (256, 256, 2, 'none'),
                                                                   (512, 256, 20, 'none'), (768, 256, 4, 'sigmoid')]
    assert [o.dst.binding for o in projs] == [0, 1, 1, 3]
    keys = list(g1.spec)
    assert keys.index('core.uncertainty_head.block.0.weight') > keys.index('core.fourier_head.block.4.bias')
    assert keys.index('core.uncertainty_head.block.4.bias') < keys.index('core.refinement_head.block.0.weight')
    assert g1.spec['core.refinement_head.block.4.weight'][0] == (12, 64, 1, 1)


def test_every_architecture_is_tensor_core_eligible():
    """All convolutions of all 19 architectures satisfy the tcgen05 engine's layout constraints (64-multiples after the
    im2col stem), which is what lets the fp16x3 parity engine run them."""
    for arch in G.ARCHS:
        g = G.trace(arch, 1, 64, 64, stem_im2col=True)
        for op in g.ops:
            if op.kind == 'conv':
                assert PL.engine_for(op, True, op.src.c) == _lib.ENGINE_TCGEN05, (arch, op.name)


def test_f16f8_packing_reproduces_fp32_products():
    """CPU emulation of the 2-pass engine's arithmetic from the PACKED operands (celldetection_b200.models.plan.pack_f16f8
    for the weights, ops.conv._f16f8_nhwc for the activations, i.e. the byte layouts of cpn_b200.h CPN_DT_F16F8):
    acc = A_hi x fp16(S W_hi) + (lo8 | hi8) x (W_hi8 ; W_lo8), out = acc * acc_scale.  Checks the chunk pairing and that
    both halves of the 8-bit product carry the main pass's scale; the result must sit ~2^4 closer to the fp32 product
    than single-pass fp16."""
    from celldetection_b200.models.plan import pack_f16f8
    from celldetection_b200.ops.conv import _f16f8_nhwc
    g = torch.Generator().manual_seed(0)
    for amp in (0.02, 1.0, 37.):
        k, cout, npx = 128, 64, 50
        w = torch.randn(1, cout, k, generator=g) * amp
        x = torch.randn(npx, k, 1, 1, generator=g) * 3.
        blob, acc_scale = pack_f16f8(w)
        w8 = blob[0, :, :2 * k].reshape(cout, k // 32, 2, 32)
        wh8 = w8[:, :, 0].reshape(cout, k).contiguous().view(torch.float8_e4m3fn).double()
        wl8 = w8[:, :, 1].reshape(cout, k).contiguous().view(torch.float8_e4m3fn).double()
        w16 = blob[0, :, 2 * k:].contiguous().view(torch.float16).double()
        px = _f16f8_nhwc(x, k)[:, 0, 0]                       # [npx, 2k] fp16-typed
        a_hi = px[:, :k].double()
        a8 = px[:, k:].contiguous().view(torch.uint8).reshape(npx, k // 32, 2, 32)
        a_lo8 = a8[:, :, 0].reshape(npx, k).contiguous().view(torch.float8_e4m3fn).double()
        a_hi8 = a8[:, :, 1].reshape(npx, k).contiguous().view(torch.float8_e4m3fn).double()
        acc = a_hi @ w16.T + a_lo8 @ wh8.T + a_hi8 @ wl8.T
        got = acc * acc_scale
        want = x[:, :, 0, 0].double() @ w[0].double().T
        single = x[:, :, 0, 0].half().double() @ w[0].half().double().T
        e2 = float((got - want).abs().max() / want.abs().max())
        e1 = float((single - want).abs().max() / want.abs().max())
        assert e2 < 3e-5 and e2 < e1 / 8, (amp, e1, e2)
        assert float(w16.abs().max()) < 65504 and float(wh8.abs().max()) <= 448


def test_constructor_options_are_honoured_or_rejected():
    """Reference constructor options (models/cpn.py:288-321, CPNCore :126-149) are either honoured -- the shape-changing
    head options, visible in the state_dict shapes and the traced plan -- or rejected; none is silently swallowed."""
    m = cd.models.CpnResNet18FPN(3, kernel_size_score=3, kernel_size_refinement=5, contour_head_channels=128,
                                 refinement_head_channels=64, contour_head_stride=2, refinement_head_stride=2,
                                 backbone_kwargs=dict(fpn_channels=128, pretrained=False), pretrained=False,
                                 uncertainty_factor=7., order_weights=True, contour_features='1')
    sd = m.state_dict()
    assert tuple(sd['core.score_head.block.0.weight'].shape) == (128, 128, 3, 3)
    assert tuple(sd['core.fourier_head.block.0.weight'].shape) == (128, 128, 7, 7)
    assert tuple(sd['core.refinement_head.block.0.weight'].shape) == (64, 128, 5, 5)
    assert tuple(sd['core.backbone.fpn.inner_blocks.0.0.weight'].shape) == (128, 64, 1, 1)
    g = G.trace(m.arch, 1, 128, 128, **m._variant())
    convs = {o.name: o for o in g.ops if o.kind == 'conv'}
    assert convs['heads.block.0.k7'].dst.c == 256 and convs['heads.block.0.k7'].stride == 2      # location + fourier merged
    assert convs['heads.block.0.k3'].dst.c == 128 and convs['heads.block.0.k3'].k == 3
    assert g.head_hw == (16, 16) and g.ref_hw == (64, 64)
    assert m.hparams['kernel_size_score'] == 3 and m.hparams['backbone_kwargs']['fpn_channels'] == 128
    for bad, exc in ((dict(contour_features=['0', '1']), NotImplementedError), (dict(foo=1), TypeError),
                     (dict(backbone_kwargs=dict(inputs_mean=0.5)), NotImplementedError),
                     (dict(refinement_interpolation='nearest'), NotImplementedError),
                     (dict(head_activation='gelu'), NotImplementedError), (dict(contour_head_stride=4), NotImplementedError),
                     (dict(backbone_kwargs=dict(fpn_channels=128)), ValueError)):
        with pytest.raises(exc):
            cd.models.CpnU22(3, **bad)
    from celldetection_b200.inference import _parse_model_parameters
    assert _parse_model_parameters('nms_thresh=0.3, certainty_thresh=None,refinement=False,tag=a=b') == [
        ('nms_thresh', 0.3), ('certainty_thresh', None), ('refinement', False), ('tag', 'a=b')]


def test_up2_phase_kernels_equal_upsample_then_conv():
    """CPN_CONV_UP2 (cpn_b200.h): conv3x3(pad 1) after a nearest x2 up-sampling equals four phase kernels on the low-res
    input followed by a pixel shuffle -- exactly, borders included; the tensor-core traces use it for the U-Net bridge
    block (one op fewer, same FLOP census, same parameters)."""
    import torch.nn.functional as F
    from celldetection_b200.models.plan import up2_weights
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 7, 9, generator=g, dtype=torch.float64)
    w = torch.randn(6, 5, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(6, generator=g, dtype=torch.float64)
    want = F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), w, b, padding=1)
    we, be = up2_weights(w, b)
    y = F.conv2d(x, we, be, padding=1)
    n, _, h, ww = y.shape
    got = y.reshape(n, 2, 2, 6, h, ww).permute(0, 3, 4, 1, 5, 2).reshape(n, 6, 2 * h, 2 * ww)
    assert float((got - want).abs().max()) < 1e-12
    a = G.trace('CpnResNeXt101UNet', 1, 128, 128, stem_im2col=True)
    c = G.trace('CpnResNeXt101UNet', 1, 128, 128, stem_im2col=True, fuse_up2=True)
    assert list(a.spec) == list(c.spec) and not any(o.kind == 'upsample' for o in c.ops)
    up = [o for o in c.ops if o.up2]
    # the bridge block (whole conv) + the up-sampled halves of the four decoder convs over cat(lateral, up(top))
    assert [o.name.rsplit('.', 3)[-3:] for o in up][-1] == ['layer_blocks', '0', '0'] and len(up) == 5
    for o in up[:-1]:
        lat = [q for q in c.ops if q.name == o.name[:-3]][0]
        assert o.name.endswith('.up') and o.act == 'none' and o.params.no_bias and lat.res is o.dst and lat.act == 'relu'
        assert lat.params.cin_range == (0, lat.src.c) and o.params.cin_range == (lat.src.c, lat.src.c + o.src.c)
        assert (o.dst.h, o.dst.w) == (2 * o.src.h, 2 * o.src.w) == (lat.dst.h, lat.dst.w)
    # the phase N tiles issue 4 of 9 taps: fewer executed FLOPs than the reference formulation
    assert G.conv_flops(c) < G.conv_flops(a)
    ragged = G.trace('CpnResNeXt101UNet', 1, 90, 120, stem_im2col=True, fuse_up2=True)   # non-2x levels keep up-sample + cat
    assert any(o.kind == 'upsample' for o in ragged.ops)


@pytest.mark.parametrize('k', [7, 5, 3])
def test_bilinear2_phase_weights_equal_upsample_then_conv(k):
    """``conv_kxk(interpolate(x, x2, bilinear))`` == the composed phase convolution on the low-res map (plan.bilinear2_weights)
    outside an output border of k//2 + 1 pixels, and -- for the k = 7 head -- everywhere once the two merged border strips
    ((top 4 | bottom 4) rows, (left 4 | right 4) columns of the low-res map) are recomputed by the plain path, which is what
    models/cpn.py:_assemble_refinement does on the GPU (float64, exact to rounding)."""
    torch.manual_seed(k)
    L = torch.randn(2, 6, 13, 17, dtype=torch.float64)
    w = torch.randn(5, 6, k, k, dtype=torch.float64)
    b = torch.randn(5, dtype=torch.float64)
    up = lambda t: F.interpolate(t, scale_factor=2, mode='bilinear', align_corners=False)      # noqa: E731
    ref = F.conv2d(up(L), w, b, padding=k // 2)
    wp, bp = PL.bilinear2_weights(w, b)
    D = wp.shape[-1] // 2
    assert wp.shape == (20, 6, 2 * D + 1, 2 * D + 1) and D == (k // 2 + 2) // 2
    o = F.conv2d(L, wp, bp, padding=D).reshape(2, 2, 2, 5, 13, 17).permute(0, 3, 4, 1, 5, 2).reshape(2, 5, 26, 34)
    m = k // 2 + 1
    assert float((o - ref)[..., m:-m, m:-m].abs().max()) < 1e-12
    assert float((o - ref).abs().max()) > 1e-3                      # the border really differs
    if k == 7:
        sl = 4
        tb = F.conv2d(up(torch.cat((L[:, :, :sl], L[:, :, -sl:]), 2)), w, b, padding=3)
        o[:, :, :m], o[:, :, -m:] = tb[:, :, :m], tb[:, :, -m:]
        lr = F.conv2d(up(torch.cat((L[:, :, :, :sl], L[:, :, :, -sl:]), 3)), w, b, padding=3)
        o[:, :, :, :m], o[:, :, :, -m:] = lr[:, :, :, :m], lr[:, :, :, -m:]
        assert float((o - ref).abs().max()) < 1e-12


def test_phase_refinement_trace_keeps_the_reference_flop_count_and_keys():
    """The phase-decomposed refinement head changes ops, not parameters: same state_dict spec, same FLOPs in the reference
    formulation, a 5x5 convolution with four fused projections in place of bilinear + 7x7 + projection."""
    a = G.trace('CpnResNet18FPN', 2, 128, 128, stem_im2col=True, fuse_up2=True)
    b = G.trace('CpnResNet18FPN', 2, 128, 128, stem_im2col=True, fuse_up2=True, phase_refinement=True)
    assert list(a.spec.keys()) == list(b.spec.keys())
    assert G.conv_flops(a) == G.conv_flops(b)
    assert a.ref_phase is None and b.ref_phase is not None and b.ref_hw == (64, 64)
    assert [o.kind for o in b.ops[-5:]] == ['conv', 'proj', 'proj', 'proj', 'proj'] and b.ops[-5].k == 5
    assert not any(o.kind == 'bilinear' for o in b.ops) and any(o.kind == 'bilinear' for o in a.ops)
    # not applicable: refinement features already at the input resolution (U-Net models), other kernel sizes
    assert G.trace('CpnU22', 1, 64, 64, stem_im2col=True, phase_refinement=True).ref_phase is None
    assert G.trace('CpnResNet18FPN', 1, 128, 128, stem_im2col=True, phase_refinement=True,
                   kernel_sizes=dict(refinement=5)).ref_phase is None
    s = G.trace_ring_strip(2, 8, 64, 256, 256, 2, 3.)
    assert [o.kind for o in s.ops] == ['bilinear', 'conv', 'proj'] and s.ref_hw == (16, 128)


def _interpret_trace(g, sd, x):
    """CPU interpreter of a (non-fused) trace with torch ops on the FOLDED weights the packer would upload (plan.fold_conv):
    validates the graph lowering -- op order, concatenation buffers, channel padding, BN folding, projections -- without a
    GPU.  Returns {output key: NCHW tensor}."""
    bufs = {}

    def root_buf(t):
        r, off = t.root()
        if r.id not in bufs:
            bufs[r.id] = torch.zeros(x.shape[0], r.c, r.h, r.w, dtype=torch.float64)
        return bufs[r.id], off

    def read(t):
        b, off = root_buf(t)
        return b[:, off:off + t.c]

    def write(t, v):
        b, off = root_buf(t)
        assert tuple(v.shape[1:]) == (t.c, t.h, t.w), (tuple(v.shape), (t.c, t.h, t.w))
        b[:, off:off + t.c] = v

    acts = dict(none=lambda v, s: v, relu=lambda v, s: F.relu(v), sigmoid=lambda v, s: torch.sigmoid(v),
                scaled_tanh=lambda v, s: torch.tanh(v) * s)
    for op in g.ops:
        if op.kind == 'prep':
            write(op.dst, x.double())
        elif op.kind == 'conv':
            assert op.im2col is None and op.gather is None
            w, b = PL.fold_conv(sd, op.params)
            if op.up2 or op.bilin2:        # phase convolutions on the low-res map: transformed weights, then pixel shuffle
                w, b = PL.up2_weights(w, b) if op.up2 else PL.bilinear2_weights(w, b)
                v = F.conv2d(read(op.src), w.double(), b.double(), padding=w.shape[-1] // 2)
                if op.up2:
                    n_, c4, hh, ww = v.shape
                    v = v.reshape(n_, 2, 2, c4 // 4, hh, ww).permute(0, 3, 4, 1, 5, 2).reshape(n_, c4 // 4, 2 * hh, 2 * ww)
            else:
                v = F.conv2d(read(op.src), w.double(), b.double(), stride=op.stride, padding=op.pad, groups=op.params.groups)
            if op.res is not None:
                r = read(op.res)
                if tuple(r.shape[2:]) != tuple(v.shape[2:]):      # FPN top-down path: nearest up-sampling inside the residual add
                    r = F.interpolate(r, size=v.shape[2:], mode='nearest')
                v = v + r
            write(op.dst, acts[op.act](v, op.act_scale))
        elif op.kind == 'proj':
            w, b = PL.fold_conv(sd, op.params)
            v = F.conv2d(read(op.src)[:, op.cin_off:op.cin_off + op.cin], w.double(), b.double())
            write(op.dst, acts[op.act](v, op.act_scale))
        elif op.kind == 'maxpool':
            write(op.dst, F.max_pool2d(read(op.src), op.k, op.stride, op.pad))
        elif op.kind == 'upsample':
            write(op.dst, F.interpolate(read(op.src), size=(op.dst.h, op.dst.w), mode='nearest'))
        elif op.kind == 'bilinear':
            write(op.dst, F.interpolate(read(op.src), size=(op.dst.h, op.dst.w), mode='bilinear', align_corners=False))
        else:
            raise AssertionError(op.kind)
    return {k: read(t) for k, t in g.outputs.items()}


@pytest.mark.parametrize('arch,hw', [('CpnSlimU22', (64, 80)), ('CpnU22', (48, 64)), ('CpnResUNet', (64, 64)),
                                     ('CpnResNet18FPN', (64, 96)), ('CpnResNeXt50UNet', (64, 64))])
def test_trace_interpreted_on_the_cpu_equals_the_oracle(arch, hw):
    """The lowered graph + folded (and, for CpnSlimU22, zero-padded) weights reproduce the oracle's head tensors."""
    import cpn_oracle as orc
    from celldetection_b200.utils.synth import synth_state_dict
    sd = synth_state_dict(key_spec(arch), seed=5)
    torch.manual_seed(2)
    h, w = hw
    x = torch.rand(1, 3, h, w)
    g = G.trace(arch, 1, h, w)
    got = _interpret_trace(g, sd, x)
    with torch.no_grad():
        s, l, r, f = orc.cpn_core(x, sd, arch)
    want = dict(scores=s, locations=l, fourier=f, refinement=r)
    lf = got['locfou']
    outs = dict(scores=got['scores'], locations=lf[:, :2], fourier=lf[:, 2:], refinement=got['refinement'])
    for k, v in want.items():
        a = outs[k].float()
        if tuple(a.shape[2:]) != tuple(v.shape[2:]):          # strided / low-res refinement is resized by the caller
            a = F.interpolate(a, size=v.shape[2:], mode='bilinear', align_corners=False)
        assert tuple(a.shape) == tuple(v.shape), (k, a.shape, v.shape)
        err = float((a - v).abs().max()) / max(float(v.abs().max()), 1e-12)
        assert err < 3e-4, (arch, k, err)     # float64 interpreter vs the fp32 oracle; a lowering error is O(1)
    if arch == 'CpnSlimU22':                                 # padded channel blocks stay exactly zero
        assert any(op.params.pad_out for op in g.ops if op.kind == 'conv')


@pytest.mark.parametrize('arch,hw', [('CpnResNeXt50UNet', (64, 64)), ('CpnU22', (64, 96)), ('CpnResNet18FPN', (64, 96))])
def test_fused_trace_interpreted_on_the_cpu_equals_the_oracle(arch, hw):
    """The tensor-core engines' graph rewrites with their transformed weights -- bridge block and decoder convolutions as
    phase convolutions on the low-res map (up2_weights, the split over cat(lateral, up(top))), the phase-decomposed refinement
    head (bilinear2_weights; compared outside the 4-pixel border the border strips recompute) -- against the oracle."""
    import cpn_oracle as orc
    from celldetection_b200.utils.synth import synth_state_dict
    sd = synth_state_dict(key_spec(arch), seed=6)
    torch.manual_seed(4)
    h, w = hw
    x = torch.rand(1, 3, h, w)
    g = G.trace(arch, 1, h, w, fuse_up2=True, phase_refinement=True)
    assert any(op.up2 for op in g.ops if op.kind == 'conv') or arch.endswith('FPN')
    got = _interpret_trace(g, sd, x)
    with torch.no_grad():
        s, l, r, f = orc.cpn_core(x, sd, arch)
    lf = got['locfou']
    ref = got['refinement']
    if g.ref_phase is not None:            # phase-packed records [1, 4 * c2, h/2, w/2] -> [1, c2, h, w]
        c2 = g.ref_phase['c2']
        ref = ref.reshape(1, 2, 2, c2, h // 2, w // 2).permute(0, 3, 4, 1, 5, 2).reshape(1, c2, h, w)
        ref, r = ref[..., 4:-4, 4:-4], r[..., 4:-4, 4:-4]
    for k, a, v in (('scores', got['scores'], s), ('locations', lf[:, :2], l), ('fourier', lf[:, 2:], f), ('refinement', ref, r)):
        assert tuple(a.shape) == tuple(v.shape), (k, a.shape, v.shape)
        err = float((a.float() - v).abs().max()) / max(float(v.abs().max()), 1e-12)
        assert err < 3e-4, (arch, k, err)
    assert (g.ref_phase is not None) == arch.endswith('FPN')


def test_work_items_are_sharded_without_loss_or_overlap():
    """Tiles x test-time repetitions dealt round-robin to the ranks (inference.shard_items): every item exactly once, the
    union in key order is the single-process sequence, tile / repetition recovered like TileLoader.__getitem__."""
    from celldetection_b200.inference import shard_items
    tiles = [0, 1, 2, 5, 6, 9]                      # tiles 3, 4, 7, 8 were skipped by an empty mask crop
    for reps in (1, 3):
        single = shard_items(tiles, reps, 0, 1)
        assert single == sorted(single) and len(single) == len(tiles) * reps
        assert [(i // reps, i % reps) for i in single] == [(t, r) for t in tiles for r in range(reps)]
        for world in (2, 3, 8, 32):
            parts = [shard_items(tiles, reps, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == single
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
