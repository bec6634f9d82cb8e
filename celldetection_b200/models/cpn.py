"""``cd.models.CPN`` drop-in for the inference path, executing on the C ABI's sm_100a kernels.

Mirrors the public surface of /root/reference/celldetection/models/cpn.py (``CPN`` :287, ``CpnU22`` :772,
``CpnResNeXt101UNet`` :930, ``CpnResNet18FPN`` :1250): constructor arguments, mutable attributes read on every call
(``order, nms_thresh, samples, score_thresh, refinement_iterations, certainty_thresh, uncertainty_nms``, :368-380),
``nn.Module`` semantics with the reference's ``state_dict`` key layout, and ``model(inputs, targets=None, nms=True,
**kwargs) -> OrderedDict`` of per-image lists (:710-734).  The inference variants are covered: ``classes > 2``
(softmax / argmax scoring, :583-585), ``uncertainty_head`` with ``certainty_thresh`` / ``uncertainty_nms`` (:617-618,
:723-726) and ``refinement_buckets > 1`` (:73-82).  Training (``compute_loss``) raises ``NotImplementedError``.

Nothing here computes on the CPU or with PyTorch operators: modules only *hold* parameters; ``forward`` lowers to
``cpn_plan_forward`` + the post-head C entry points.  PyTorch supplies device memory, the stream and list plumbing.
"""
from collections import OrderedDict
import ctypes
import math
import os

import torch
import torch.nn as nn
from torch import Tensor

from .. import _lib as L
from ..ops import cpn as O
from .graph import trace, trace_sparse_heads, trace_ring_strip, ARCHS, HEAD_KERNEL_KEYS
from .plan import Plan, WeightPack, SPLIT_NONE, SPLIT_X3, SPLIT_F8


PRECISIONS = ('fp16f8', 'fp16', 'fp16x3', 'fp32')
_SPLIT = dict(fp16f8=SPLIT_F8, fp16=SPLIT_NONE, fp16x3=SPLIT_X3)


class _Node(nn.Module):
    """Parameter container mirroring one node of the reference's module tree (never called)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError('celldetection_b200 modules hold parameters only; call the CPN model itself.')


def _get_node(root: nn.Module, path):
    m = root
    for p in path:
        if p not in m._modules:
            m.add_module(p, _Node())
        m = m._modules[p]
    return m


# Constructor options of the reference (models/cpn.py:288-321, CPNCore :126-149) that do not change the inference
# arithmetic when left at these values; any other value is rejected (never silently ignored).
_NOOP_DEFAULTS = dict(contour_features='1', location_features='1', uncertainty_features='1', score_features='1',
                      refinement_features='0', order_weights=True, refinement_interpolation='bilinear',
                      head_activation='relu', head_activation_score='relu', head_activation_location='relu',
                      head_activation_fourier='relu', head_activation_uncertainty='relu',
                      head_activation_refinement='relu', fuse_kwargs={}, encoder_channels=None, pretrained=False,
                      encoder_prefix='encoder.')
_TRAINING_ONLY = ('uncertainty_factor', 'score_target_dtype')          # used by compute_loss only
# backbone_kwargs of the ResNet U-Net / FPN backbones (unet.py:251-272, fpn.py:137-170) that are the identity here
_BACKBONE_NOOPS = dict(pretrained=False, normalize=True, inputs_mean=0., inputs_std=1., assert_range=(0., 1.),
                       interpolate='nearest', nd=2, block=None, block_kwargs=None, final_activation=None)


def _same(a, b):
    if isinstance(a, (list, tuple)) or isinstance(b, (list, tuple)):
        return a is not None and b is not None and list(a) == list(b)
    return a == b or (a is None and b in (False, {}, None)) or (b is None and a in (False, {}, None))


class CPN(nn.Module):
    def __init__(self, backbone: str, in_channels: int = 3, order: int = 5, nms_thresh: float = .2,
                 score_thresh: float = .9, certainty_thresh: float = None, samples: int = 32, classes: int = 2,
                 refinement: bool = True, refinement_iterations: int = 4, refinement_margin: float = 3.,
                 refinement_buckets: int = 1, uncertainty_head=False, uncertainty_nms=False,
                 contour_head_channels: int = None, contour_head_stride: int = 1, refinement_head_channels: int = None,
                 refinement_head_stride: int = 1, refinement_full_res: bool = True, backbone_kwargs: dict = None,
                 precision: str = 'fp16f8', **kwargs):
        """Contour Proposal Network (inference).

        Args:
            backbone: architecture name, one of ``models.graph.ARCHS`` -- ``CpnU22`` and ``Cpn<ResNet-family><UNet|FPN>``
                (the reference takes a backbone *module*; here the backbone is part of the compiled plan).
            order, nms_thresh, score_thresh, samples, classes, refinement, refinement_iterations, refinement_margin,
                refinement_buckets: as in the reference (models/cpn.py:288-321).
            precision: ``'fp16f8'`` (default) -- tcgen05 tensor-core engine that meets the reference's fp32 results to
                1e-3: fp16 activations / weights plus e4m3 copies of their rounding residuals, one ``kind::f16`` pass and
                one ``kind::f8f6f4`` correction pass per K block into the same fp32 accumulator (2 pass-equivalents);
                ``'fp16'`` -- the single-pass engine (fp16 operands, fp32 accumulation): fastest, head tensors only
                within ~1e-2 of the reference (opt-in); ``'fp16x3'`` -- activations and weights as fp16 (hi, lo) pairs,
                three fp16 passes (hi*hi + lo*hi + hi*lo): fp32-level accuracy at a third of the tensor throughput;
                ``'fp32'`` -- strict CUDA-core fp32 engine.
        """
        super().__init__()
        if backbone not in ARCHS:
            raise ValueError(f'Unknown backbone/architecture {backbone!r}; available: {ARCHS}')
        if classes < 1 or classes > 32:
            raise ValueError('classes must be in [1, 32]')
        if refinement_buckets < 1 or refinement_buckets > 12:
            raise ValueError('refinement_buckets must be in [1, 12]')
        if precision not in PRECISIONS:
            raise ValueError(f'precision must be one of {PRECISIONS}')
        # ---- shape-changing head options (models/cpn.py:177-234): honoured; everything else must be a no-op ----
        extra = dict(kwargs)
        self.kernel_sizes = {k: int(kwargs.pop(f'kernel_size_{k}', 7)) for k in HEAD_KERNEL_KEYS}
        for k in _TRAINING_ONLY:
            kwargs.pop(k, None)
        for k, dflt in _NOOP_DEFAULTS.items():
            if k in kwargs and not _same(kwargs[k], dflt):
                raise NotImplementedError(f'{k}={kwargs[k]!r} is outside the accelerated path (only {dflt!r} is supported)')
            kwargs.pop(k, None)
        if kwargs:
            raise TypeError(f'unsupported CPN options: {sorted(kwargs)} (they would change the network and are not '
                            f'implemented here)')
        bkw = dict(backbone_kwargs or {})
        self.fpn_channels = int(bkw.pop('fpn_channels', 256))
        if self.fpn_channels != 256 and not backbone.endswith('FPN'):
            raise ValueError('fpn_channels applies to the FPN backbones only')
        for k, v in bkw.items():
            if k not in _BACKBONE_NOOPS or not _same(v, _BACKBONE_NOOPS[k]):
                raise NotImplementedError(f'backbone_kwargs[{k!r}]={v!r} is outside the accelerated path')
        for nm, v in (('contour_head_stride', contour_head_stride), ('refinement_head_stride', refinement_head_stride)):
            if int(v) not in (1, 2):
                raise NotImplementedError(f'{nm}={v}: strides 1 and 2 are supported')
        self.contour_head_channels = None if not contour_head_channels else int(contour_head_channels)
        self.refinement_head_channels = None if not refinement_head_channels else int(refinement_head_channels)
        self.contour_head_stride, self.refinement_head_stride = int(contour_head_stride), int(refinement_head_stride)
        self.refinement_full_res = bool(refinement_full_res)
        self.arch = backbone
        self.in_channels = in_channels
        self.order = order
        self.core_order = order
        self.nms_thresh = nms_thresh
        self.samples = samples
        self.score_thresh = score_thresh
        self.score_channels = 1 if classes in (1, 2) else classes          # models/cpn.py:372
        self.refinement = refinement
        self.refinement_iterations = refinement_iterations
        self.refinement_margin = refinement_margin
        self.refinement_buckets = int(refinement_buckets)
        self.uncertainty_head = bool(uncertainty_head)
        self.certainty_thresh = certainty_thresh
        self.uncertainty_nms = uncertainty_nms
        self.precision = precision
        # Location / fourier heads only where the score selects a proposal (the reference computes them on every pixel,
        # models/cpn.py:253-263, and reads them at the selected pixels only, :620-623): same outputs, ~2/3 of the head
        # convolution's work gone.  ``core_forward`` (raw head tensors) always runs the dense plan.
        self.sparse_heads = True
        # Refinement head on bilinearly x2 up-sampled features (the FPN models, models/cpn.py:274-279) as four phase
        # convolutions on the low-res features (25 instead of 49 taps per output, no full-resolution feature tensor); the
        # 4-pixel image border is recomputed by the plain path on two strips (see ``_assemble_refinement``).
        self.phase_refinement = os.environ.get('CPN_REF_PHASE', '1') != '0'
        self.cuda_graph = False     # True: replay the backbone + heads as one CUDA graph (static buffers; see Plan.forward_graph)
        self.hparams = dict(in_channels=in_channels, order=order, nms_thresh=nms_thresh, score_thresh=score_thresh,
                            samples=samples, classes=classes, refinement=refinement,
                            refinement_iterations=refinement_iterations, refinement_margin=refinement_margin,
                            refinement_buckets=refinement_buckets, uncertainty_head=uncertainty_head,
                            uncertainty_nms=uncertainty_nms, certainty_thresh=certainty_thresh,
                            contour_head_channels=contour_head_channels, contour_head_stride=contour_head_stride,
                            refinement_head_channels=refinement_head_channels,
                            refinement_head_stride=refinement_head_stride, refinement_full_res=refinement_full_res,
                            **extra)
        if backbone_kwargs:
            self.hparams['backbone_kwargs'] = dict(backbone_kwargs)
        # ---- parameters / buffers with the reference's state_dict keys ----
        g = trace(backbone, 1, 64, 64, in_channels=in_channels, order=order, refinement_margin=refinement_margin,
                  **self._variant())
        self._spec = g.spec
        gen = torch.Generator().manual_seed(torch.initial_seed() & 0x7fffffff)
        for key, (shape, role) in g.spec.items():
            *path, leaf = key.split('.')
            node = _get_node(self, path)
            if role == 'order_weights':
                x = torch.arange(order).float()
                spread = max(order - 1, 1)
                node.register_buffer(leaf, (1 + 4 * (1 - (x / spread).clamp(0., 1.)) ** 2)[:, None])  # ops/cpn.py:230-235
            elif role == 'conv_w':
                w = torch.empty(shape)
                fan_in = shape[1] * shape[2] * shape[3]
                a = 1. if ('.unet.' in key or '.fpn.' in key) else math.sqrt(5.)  # unet.py:171-176, fpn.py:125-129
                bound = math.sqrt(6. / ((1 + a * a) * fan_in))
                w.uniform_(-bound, bound, generator=gen)
                node.register_parameter(leaf, nn.Parameter(w))
            elif role == 'conv_b':
                wshape = g.spec[key[:-len('bias')] + 'weight'][0]
                fan_in = wshape[1] * wshape[2] * wshape[3]
                b = torch.zeros(shape)
                if not ('.unet.' in key or '.fpn.' in key):
                    bound = 1. / math.sqrt(fan_in)
                    b.uniform_(-bound, bound, generator=gen)
                node.register_parameter(leaf, nn.Parameter(b))
            elif role in ('bn_w', 'bn_w_res'):
                node.register_parameter(leaf, nn.Parameter(torch.ones(shape)))
            elif role == 'bn_b':
                node.register_parameter(leaf, nn.Parameter(torch.zeros(shape)))
            elif role == 'bn_rm':
                node.register_buffer(leaf, torch.zeros(shape))
            elif role == 'bn_rv':
                node.register_buffer(leaf, torch.ones(shape))
            elif role == 'bn_nbt':
                node.register_buffer(leaf, torch.tensor(0, dtype=torch.long))
            else:  # pragma: no cover
                raise AssertionError(role)
        self.eval()
        self._packs = {}
        self._plans = {}
        self._sparse_plans = {}
        self._ws = {}

    def _variant(self):
        return dict(score_channels=self.score_channels, refinement_buckets=self.refinement_buckets,
                    uncertainty_head=self.uncertainty_head, kernel_sizes=self.kernel_sizes,
                    contour_head_channels=self.contour_head_channels,
                    refinement_head_channels=self.refinement_head_channels,
                    contour_head_stride=self.contour_head_stride, refinement_head_stride=self.refinement_head_stride,
                    refinement_full_res=self.refinement_full_res, fpn_channels=self.fpn_channels)

    # ---- plan management ------------------------------------------------------------------------------------------
    def invalidate_plans(self):
        """Drop packed weights / compiled plans (call after modifying parameters in place)."""
        self._packs.clear()
        self._plans.clear()
        self._sparse_plans.clear()
        self._ws.clear()

    def load_state_dict(self, state_dict, strict=True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self.invalidate_plans()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        if hasattr(self, '_packs'):
            self.invalidate_plans()
        return r

    @property
    def device(self):
        return self.order_weights.device

    SPARSE_ROWS = 32768      # proposals per sparse-heads launch (the gathered matrix holds 50 KB per row at k = 7, C = 256)

    def _sparse_plan(self, rows) -> Plan:
        """Plan of the location + fourier heads on `rows` gathered patches (graph.trace_sparse_heads)."""
        key = (rows, self.precision)
        plan = self._sparse_plans.get(key)
        if plan is None:
            ref = self._sparse_ref
            g = trace_sparse_heads(rows, ref['head_c'], ref['head_mid'], ref['k'], self.core_order)
            pk = ('sparse', self.precision)
            pack = self._packs.get(pk)
            if pack is None:
                with torch.no_grad():
                    pack = WeightPack(g, self.state_dict(), True, self.device, split=_SPLIT[self.precision])
                self._packs[pk] = pack
            plan = Plan(g, pack, True, self.device, split=_SPLIT[self.precision])
            self._sparse_plans[key] = plan
        return plan

    def _sparse_records(self, plan: Plan, idx: Tensor, P: int) -> Tensor:
        """Head records [>= P, 2 + 4*order_core] (loc_x, loc_y, fourier...) of the P proposals `idx` of `plan`'s head
        feature map: cpn_gather_patches + the sparse-heads plan, at most SPARSE_ROWS proposals per launch."""
        lib = L.load()
        g = plan.g
        self._sparse_ref = dict(head_c=g.head_feat.c, head_mid=g.head_mid, k=g.head_k)
        width = 2 + 4 * int(self.core_order)
        chunks, start = [], 0
        while start < P:
            n = min(self.SPARSE_ROWS, P - start)
            rows = 128
            while rows < n:
                rows *= 2
            chunks.append((start, n, rows))
            start += n
        rec = torch.empty((chunks[-1][0] + chunks[-1][2], width), dtype=torch.float32, device=idx.device)
        src_view = plan.view_of(g.head_feat)
        for start, n, rows in chunks:                      # in order: a chunk's padding rows are overwritten by the next
            sp = self._sparse_plan(rows)
            dst_view = sp.view_of(sp.g.input_tensor)
            L.check(lib.cpn_gather_patches(L.ptr(plan.arena), ctypes.byref(src_view), L.ptr(idx[start:]), n, g.head_k,
                                           L.ptr(sp.arena), ctypes.byref(dst_view), L.stream_ptr()), 'gather_patches')
            if os.environ.get('CPN_SPARSE_TRIM', '1') != '0':
                sp.set_active_rows(n)        # the row matrix is sized for `rows` (a power of two), only n are proposals
            sp.forward(None, L.IN_F32_NCHW, [rec[start:]])
        self.last_sparse_rows = sum(c[2] for c in chunks)
        return rec

    def _strip_plan(self, n, hs, ws, ref) -> Plan:
        """Plan of the refinement head on a cropped [n, hs, ws] strip of the low-res features (graph.trace_ring_strip)."""
        key = ('ring', n, hs, ws, self.precision)
        plan = self._sparse_plans.get(key)
        if plan is None:
            g = trace_ring_strip(n, hs, ws, ref['c'], ref['mid'], ref['c2'], ref['margin'], ref['k'])
            pk = ('ring', self.precision)
            pack = self._packs.get(pk)
            if pack is None:            # both strip orientations have the same op list: one packed copy of the head
                with torch.no_grad():
                    pack = WeightPack(g, self.state_dict(), True, self.device, split=_SPLIT[self.precision])
                self._packs[pk] = pack
            plan = Plan(g, pack, True, self.device, split=_SPLIT[self.precision])
            self._sparse_plans[key] = plan
        return plan

    def _assemble_refinement(self, plan: Plan, rec: Tensor, n, h, w) -> Tensor:
        """Full-resolution refinement tensor [n, h, w, 2B] of a ``phase_refinement`` plan: the phase-packed records are
        shuffled to full resolution; the image border (4 pixels), where bilinear clamping and the convolution's zero padding
        break the phase identity, is recomputed by the plain path on two strips of the low-res features -- (top 4 rows |
        bottom 4 rows) and (left 4 columns | right 4 columns); the seam inside a strip only touches outputs that are not
        used -- and pasted over it."""
        lib = L.load()
        ref = plan.g.ref_phase
        feat = ref['feat']
        hl, wl, c2 = int(feat.h), int(feat.w), int(ref['c2'])
        st = L.stream_ptr()
        out = torch.empty((n, h, w, c2), dtype=torch.float32, device=rec.device)
        L.check(lib.cpn_unshuffle2(L.ptr(rec), n, hl, wl, c2, L.ptr(out), st), 'unshuffle2')
        sv = plan.view_of(feat)
        px_bytes = int(sv.pitch) * int(plan.views['es'])
        src = ctypes.c_void_p(plan.arena.data_ptr() + int(sv.offset))
        m, sl = 4, 4                    # border width in output pixels (k//2 + 1), low-res rows / columns per side
        for rows, hs, ws in ((True, 2 * sl, wl), (False, hl, 2 * sl)):   # (top | bottom) rows, (left | right) columns
            sp = self._strip_plan(n, hs, ws, ref)
            dv = sp.view_of(sp.g.input_tensor)
            assert int(dv.pitch) * int(sp.views['es']) == px_bytes
            dst = ctypes.c_void_p(sp.arena.data_ptr() + int(dv.offset))
            for part in range(2):       # crop the two sides into the strip's input
                ys, xs = ((0 if part == 0 else hl - sl), 0) if rows else (0, (0 if part == 0 else wl - sl))
                yd, xd = (part * sl, 0) if rows else (0, part * sl)
                L.check(lib.cpn_copy_window(src, dst, n, hl, wl, hs, ws, px_bytes, ys, xs, yd, xd,
                                            sl if rows else hl, wl if rows else sl, st), 'copy_window')
            strip = sp.forward(None, L.IN_F32_NCHW)[0]              # [n, 2 hs, 2 ws, c2]
            for part in range(2):       # paste the m valid output rows / columns of each side
                if rows:
                    args = (2 * hs, 2 * ws, h, w, c2 * 4, (0 if part == 0 else 2 * hs - m), 0, (0 if part == 0 else h - m), 0, m, w)
                else:
                    args = (2 * hs, 2 * ws, h, w, c2 * 4, 0, (0 if part == 0 else 2 * ws - m), 0, (0 if part == 0 else w - m), h, m)
                L.check(lib.cpn_copy_window(L.ptr(strip), L.ptr(out), n, *args, st), 'copy_window')
        return out

    def _plan(self, n, h, w, dense=False) -> Plan:
        dev = self.device
        if dev.type != 'cuda':
            raise RuntimeError('celldetection_b200.CPN runs on CUDA (sm_100a) only: move the model with .cuda(). '
                               'There is no CPU fallback.')
        fast = self.precision in _SPLIT
        split = _SPLIT.get(self.precision, SPLIT_NONE)
        sparse = bool(self.sparse_heads) and fast and not dense
        phase = bool(self.phase_refinement) and fast
        key = (n, h, w, self.precision, sparse, phase)
        plan = self._plans.get(key)
        if plan is None:
            g = trace(self.arch, n, h, w, in_channels=self.in_channels, order=self.core_order,
                      refinement_margin=self.refinement_margin, stem_im2col=fast,
                      fuse_up2=fast and os.environ.get('CPN_UP2', '1') != '0', sparse_heads=sparse,
                      phase_refinement=phase, **self._variant())
            pk = (self.precision, bool(g.sparse), g.ref_phase is not None)
            pack = self._packs.get(pk)
            if pack is None:
                with torch.no_grad():
                    pack = WeightPack(g, self.state_dict(), fast, dev, split=split)
                self._packs[pk] = pack
            if len(self._plans) >= 8:
                self._plans.pop(next(iter(self._plans)))
            plan = Plan(g, pack, fast, dev, split=split)
            self._plans[key] = plan
        return plan

    # ---- raw head tensors (debug / parity) ------------------------------------------------------------------------
    def core_forward(self, inputs: Tensor):
        """Raw head tensors in the reference's layout: scores [N,C,h,w], locations [N,2,h,w], refinement [N,2B,H,W],
        fourier [N,4*order,h,w] (+ uncertainty [N,4,h,w] with an uncertainty head) (CPNCore.forward, cpn.py:238-283)."""
        plan, outs, _ = self._run_plan(inputs, dense=True)
        if self.cuda_graph:
            outs = [o.clone() for o in outs]       # the graph's static outputs are overwritten by the next call
        sc, lf, rf = outs[:3]
        if int(plan.flags[0].item()) & 1:
            raise AssertionError('Inputs should be in interval (0.0, 1.0)')
        out = OrderedDict(scores=sc[:, None] if sc.dim() == 3 else sc.permute(0, 3, 1, 2),
                          locations=lf[..., :2].permute(0, 3, 1, 2), refinement=rf.permute(0, 3, 1, 2),
                          fourier=lf[..., 2:].permute(0, 3, 1, 2))
        if len(outs) > 3:                      # key present only for models with an uncertainty head
            out['uncertainty'] = outs[3].permute(0, 3, 1, 2)
        return out

    # ---- post-head chain on head tensors --------------------------------------------------------------------------
    def post_flat(self, scores: Tensor, locfou: Tensor, refinement: Tensor, original_size, nms=True, offsets=None,
                  scores_lower_bound=None, scores_upper_bound=None, flags: Tensor = None, uncertainty: Tensor = None,
                  plan: Plan = None, after_count=None):
        """models/cpn.py:575-734 on device tensors: scores [N,h,w] logits ([N,h,w,C] for classes > 2), locfou
        [N,h,w,2+4*order_core] records, refinement [N,H,W,2*buckets] (or None), uncertainty [N,h,w,4] (or None).
        Returns (flat dict of concatenated tensors, rows per image)."""
        lib = L.load()
        dev = scores.device
        n, h, w = scores.shape[:3]
        channels = 1 if scores.dim() == 3 else int(scores.shape[3])
        H, W = original_size
        order = min(int(self.order), int(self.core_order))
        samples = int(self.samples)
        pixels = n * h * w
        st = L.stream_ptr()
        lo = up = None
        if scores_upper_bound is not None or scores_lower_bound is not None:
            def prep(bnd):   # _apply_score_bounds / _equal_size (cpn.py:109-123): bilinear resize to the head resolution
                if bnd is None:
                    return None
                assert bnd.dtype.is_floating_point, 'score bounds must be floating point'
                bnd = bnd.to(dev).float().contiguous()
                assert bnd.dim() == 4 and bnd.shape[0] == n and bnd.shape[1] == 1, 'score bounds must be [N,1,h,w]'
                if tuple(bnd.shape[2:]) != (h, w):
                    out = torch.empty((n, h, w), dtype=torch.float32, device=dev)
                    L.check(lib.cpn_resize_bilinear(L.ptr(bnd), n, int(bnd.shape[2]), int(bnd.shape[3]), 1, L.ptr(out),
                                                    h, w, st), 'resize_bilinear')
                    return out
                return bnd.reshape(n, h, w)
            lo, up = prep(scores_lower_bound), prep(scores_upper_bound)
        ws_key = (pixels, str(dev))
        ws = self._ws.get(ws_key)
        if ws is None:
            ws = torch.empty((int(lib.cpn_select_workspace_bytes(pixels)),), dtype=torch.uint8, device=dev)
            self._ws = {ws_key: ws}
        meta = torch.zeros((2,), dtype=torch.int64, device=dev)
        use_cert = self.certainty_thresh is not None and uncertainty is not None          # cpn.py:617-618
        sel = L.SelectParams(L.addr(scores), L.addr(lo), L.addr(up), L.addr(uncertainty), channels, int(use_cert),
                             float(self.score_thresh),
                             float(torch.tensor(1 - self.certainty_thresh, dtype=torch.float32)) if use_cert else 0.)
        L.check(lib.cpn_select_count_ex(sel, pixels, L.ptr(ws), L.ptr(meta), st), 'select_count')
        if flags is not None:
            meta[1:2].copy_(flags[:1].to(torch.int64))
        # host sync #1 (the reference syncs at torch.where): the count travels to pinned memory asynchronously; whatever
        # `after_count` enqueues (the rest of a staged plan: full-resolution branch + refinement head) keeps the GPU busy
        # while the host waits for the count and queues the proposal-dependent kernels behind it
        host_meta = self.__dict__.get('_meta_host')
        if host_meta is None:
            host_meta = self.__dict__['_meta_host'] = torch.empty((2,), dtype=torch.int64).pin_memory()
        host_meta.copy_(meta, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record()
        if after_count is not None:
            late = after_count()
            if late is not None:
                refinement = late
        ready.synchronize()
        total, flag = host_meta.tolist()
        if flag & 1:
            raise AssertionError('Inputs should be in interval (0.0, 1.0)')
        P = int(total)
        idx = torch.empty((max(P, 1),), dtype=torch.int32, device=dev)
        sel_scores = torch.empty((max(P, 1),), dtype=torch.float32, device=dev)
        classes = torch.empty((max(P, 1),), dtype=torch.long, device=dev)
        seg = torch.zeros((n + 1,), dtype=torch.int32, device=dev)
        L.check(lib.cpn_select_write_ex(sel, n, h * w, L.ptr(ws), L.ptr(idx), L.ptr(sel_scores), L.ptr(classes),
                                        max(P, 1), L.ptr(seg), st), 'select_write')
        contours = torch.empty((P, samples, 2), dtype=torch.float32, device=dev)
        proposals = torch.empty((P, samples, 2), dtype=torch.float32, device=dev)
        boxes = torch.empty((P, 4), dtype=torch.float32, device=dev)
        locations = torch.empty((P, 2), dtype=torch.float32, device=dev)
        fourier = torch.empty((P, order, 4), dtype=torch.float32, device=dev)
        use_ref = bool(self.refinement) and refinement is not None and int(self.refinement_iterations) > 0
        buckets = int(refinement.shape[-1]) // 2 if refinement is not None else 1
        off = None
        if offsets is not None:
            off = torch.as_tensor(offsets).to(device=dev, dtype=torch.float32).reshape(n, 2).contiguous()
        if P > 0:
            trig = O.trig_table(order, samples, dev)
            bidx, bw = O.bucket_table(samples, buckets, dev) if (use_ref and buckets > 1) else (None, None)
            by_row = locfou is None        # sparse heads: the records are computed now, for the selected pixels only
            if by_row:
                if plan is None or not getattr(plan.g, 'sparse', False):
                    raise ValueError('post_flat needs the dense locfou tensor or the sparse-heads plan that produced the scores')
                locfou = self._sparse_records(plan, idx, P)
            L.check(lib.cpn_decode_refine_rows(
                L.ptr(idx), P, L.ptr(locfou), int(by_row), int(self.core_order), order, n, h, w, H, W, L.ptr(trig), samples,
                L.ptr(refinement) if use_ref else None, int(self.refinement_iterations) if use_ref else 0, buckets,
                L.ptr(bidx), L.ptr(bw), L.ptr(off), L.ptr(contours), L.ptr(proposals), L.ptr(boxes), L.ptr(locations),
                L.ptr(fourier), st), 'decode_refine')
        idx, sel_scores, classes = idx[:P], sel_scores[:P], classes[:P]
        flat = OrderedDict(contours=contours, boxes=boxes, scores=sel_scores, classes=classes, locations=locations,
                           fourier=fourier, contour_proposals=proposals)
        if uncertainty is not None:                       # selected_uncertainties = uncertainty[b, :, y, x]
            unc_rows = torch.empty((P, 4), dtype=torch.float32, device=dev)
            if P > 0:
                L.check(lib.cpn_gather_rows(L.ptr(uncertainty), 16, L.ptr(idx), P, L.ptr(unc_rows), st), 'gather_rows')
            flat['box_uncertainties'] = unc_rows
        if nms:
            weights = sel_scores
            if self.uncertainty_nms and uncertainty is not None and P > 0:               # cpn.py:723-726
                weights = torch.empty_like(sel_scores)
                L.check(lib.cpn_nms_weights(L.ptr(sel_scores), L.ptr(uncertainty), L.ptr(idx), P, L.ptr(weights), st),
                        'nms_weights')
            keep, counts = O.nms_segments(boxes, weights, seg, n, float(self.nms_thresh), O.NMS_BATCH_SIZE)
            both = torch.cat((seg, counts.to(seg.dtype))).tolist()   # host sync #2 (one copy): per-image offsets + kept counts
            seg_h, counts_h = both[:n + 1], both[n + 1:]
            sel_rows = torch.cat([keep[seg_h[i]:seg_h[i] + counts_h[i]] for i in range(n)]) if P > 0 else keep[:0]
            K = int(sel_rows.numel())
            out = OrderedDict()
            for k, v in flat.items():
                row = v[0].numel() * v.element_size() if P > 0 else 0
                dst = torch.empty((K,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
                if K > 0:
                    L.check(lib.cpn_gather_rows(L.ptr(v), row, L.ptr(sel_rows), K, L.ptr(dst), st), 'gather_rows')
                out[k] = dst
            return out, counts_h
        seg_h = seg.tolist()
        return flat, [seg_h[i + 1] - seg_h[i] for i in range(n)]

    def post(self, scores, locfou, refinement, original_size, nms=True, **kw):
        """``post_flat`` + the reference's per-image list structure (resolve_batch_index, models/cpn.py:42-60)."""
        flat, sizes = self.post_flat(scores, locfou, refinement, original_size, nms=nms, **kw)
        out = OrderedDict((k, list(torch.split(v, sizes, 0))) for k, v in flat.items())
        out.setdefault('box_uncertainties', None)
        return out

    def _run_plan(self, inputs, fmt=None, dense=False, staged=False):
        if not isinstance(inputs, Tensor) or inputs.dim() != 4:
            raise ValueError('inputs must be a 4-d Tensor')
        if not inputs.is_cuda:
            raise RuntimeError('celldetection_b200.CPN expects CUDA inputs; there is no CPU fallback.')
        if fmt is None:
            fmt = L.IN_U8_NCHW if inputs.dtype == torch.uint8 else L.IN_F32_NCHW
        if fmt == L.IN_U8_NHWC:
            n, h, w, c = inputs.shape
        else:
            n, c, h, w = inputs.shape
        if c != self.in_channels:
            raise ValueError(f'expected {self.in_channels} input channels, got {c}')
        plan = self._plan(n, h, w, dense=dense)
        x = inputs.contiguous() if inputs.dtype == torch.uint8 else inputs.contiguous().float()
        plan.flags.zero_()

        def full_res(rf):
            if getattr(plan.g, 'ref_phase', None):
                return self._assemble_refinement(plan, rf, n, h, w)
            if tuple(rf.shape[1:3]) == (h, w):
                return rf
            # strided / low-res refinement head: _equal_size to the input (cpn.py:279)
            full = torch.empty((n, h, w, rf.shape[3]), dtype=torch.float32, device=rf.device)
            L.check(L.load().cpn_resize_bilinear(L.ptr(rf), n, int(rf.shape[1]), int(rf.shape[2]), int(rf.shape[3]),
                                                 L.ptr(full), h, w, L.stream_ptr()), 'resize_bilinear')
            return full

        split = getattr(plan.g, 'split_index', None)
        # staged execution is opt-in (CPN_STAGED=1): measured 343.5 / 343.7 tiles/s without vs 342.6 / 342.6 with it
        # (profiles/r02n_*): the step's tail is GPU work of the proposal kernels, not host gaps
        if staged and split and not self.cuda_graph and os.environ.get('CPN_STAGED', '0') == '1':
            # stage 1: everything up to the dense heads; stage 2 (`rest`) is enqueued by post_flat right after the
            # proposal count has been requested
            outs = plan.forward(x, fmt, None, 0, split)

            def rest():
                plan.forward(x, fmt, outs, split, None)
                return full_res(outs[2])
            return plan, outs, (h, w), rest
        outs = list(plan.forward_graph(x, fmt) if self.cuda_graph else plan.forward(x, fmt))
        outs[2] = full_res(outs[2])
        return (plan, outs, (h, w), None) if staged else (plan, outs, (h, w))

    def _post_kwargs(self, plan, outs, kwargs):
        return dict(offsets=kwargs.get('offsets'), scores_lower_bound=kwargs.get('scores_lower_bound'),
                    scores_upper_bound=kwargs.get('scores_upper_bound'), flags=plan.flags,
                    uncertainty=outs[3] if len(outs) > 3 else None, plan=plan)

    def forward_flat(self, inputs, fmt=None, nms=True, **kwargs):
        """Like ``forward`` but returns (flat dict of concatenated tensors, rows per image); accepts uint8 NHWC
        batches (``fmt=_lib.IN_U8_NHWC``) so tile crops need no host-side transpose."""
        with L.nvtx_range('cpn.plan'):
            plan, outs, hw, rest = self._run_plan(inputs, fmt, staged=True)
        with L.nvtx_range('cpn.post'):
            return self.post_flat(outs[0], outs[1], outs[2], hw, nms=nms, after_count=rest,
                                  **self._post_kwargs(plan, outs, kwargs))

    # ---- model(x) ---------------------------------------------------------------------------------------------------
    def forward(self, inputs: Tensor, targets=None, nms=True, **kwargs):
        if self.training or targets is not None:
            if self.training and targets is None:
                raise ValueError('In training mode, targets should be passed')
            raise NotImplementedError('celldetection_b200 accelerates CPN inference only (use .eval()).')
        with L.nvtx_range('cpn.plan'):
            plan, outs, hw, rest = self._run_plan(inputs, staged=True)
        with L.nvtx_range('cpn.post'):
            return self.post(outs[0], outs[1], outs[2], hw, nms=nms, after_count=rest,
                             **self._post_kwargs(plan, outs, kwargs))


def _make(arch):
    class _Cpn(CPN):
        def __init__(self, in_channels: int = 3, order: int = 5, nms_thresh: float = .2, score_thresh: float = .9,
                     samples: int = 32, classes: int = 2, refinement: bool = True, refinement_iterations: int = 4,
                     refinement_margin: float = 3., refinement_buckets: int = 1, backbone_kwargs: dict = None,
                     **kwargs):   # kwargs: uncertainty_head, uncertainty_nms, certainty_thresh, precision, the head
                                  # options of CPN.__init__ (cpn.py:288-321)
            super().__init__(arch, in_channels=in_channels, order=order, nms_thresh=nms_thresh,
                             score_thresh=score_thresh, samples=samples, classes=classes, refinement=refinement,
                             refinement_iterations=refinement_iterations, refinement_margin=refinement_margin,
                             refinement_buckets=refinement_buckets, backbone_kwargs=backbone_kwargs, **kwargs)
    _Cpn.__name__ = _Cpn.__qualname__ = arch
    return _Cpn


for _arch in ARCHS:      # CpnU22 (cpn.py:772), CpnResNeXt101UNet (:930), CpnResNet18FPN (:1250) and the rest of the family
    globals()[_arch] = _make(_arch)
__all__ = ['CPN'] + list(ARCHS)
