"""Lowering of a traced layer graph to the C ABI: weight packing (BatchNorm folding, KRSC / K-major layouts, grouped
convolutions expanded to block-diagonal 64-channel slabs), activation-arena assignment by live ranges, and the
``cpn_op_t`` array handed to ``cpn_plan_create``.

Weight packing is host work done once per (model, precision): everything is computed on the CPU and the finished blob
reaches the device with ONE host-to-device copy (no device kernels at model load).

BatchNorm folding follows SURVEY.md appendix B (eval mode): ``w' = w * gamma / sqrt(var + eps)``,
``b' = (b - mean) * gamma / sqrt(var + eps) + beta`` in fp32, then cast.
"""
import ctypes
import math
from collections import OrderedDict

import torch

from .. import _lib as L
from .graph import TT, LOp, Tracer

ALIGN = 256
BN_EPS = 1e-5
# activation storage of the tensor-core engines (cpn_b200.h): plain fp16, (hi | lo) fp16 pairs, fp16 + e4m3 corrections
SPLIT_NONE, SPLIT_X3, SPLIT_F8 = 0, 1, 2
E4M3_MAX = 448.


def _align(v, a=ALIGN):
    return (v + a - 1) // a * a


def fold_conv(sd, params):
    """Returns folded (weight [cout, cin/groups, k, k], bias [cout]) in fp32 for possibly concatenated convs."""
    ws, bs = [], []
    for wk, bk, bnk in zip(params.weight, params.bias, params.bn):
        w = sd[wk].detach().float()
        b = sd[bk].detach().float() if bk is not None else torch.zeros(w.shape[0], device=w.device)
        if bnk is not None:
            gamma, beta = sd[bnk + '.weight'].float(), sd[bnk + '.bias'].float()
            mean, var = sd[bnk + '.running_mean'].float(), sd[bnk + '.running_var'].float()
            scale = gamma / torch.sqrt(var + BN_EPS)
            w = w * scale[:, None, None, None]
            b = (b - mean) * scale + beta
        ws.append(w)
        bs.append(b)
    w, b = torch.cat(ws, 0), torch.cat(bs, 0)
    if getattr(params, 'cin_range', None) is not None:       # this op contracts over a slice of the input channels
        w = w[:, params.cin_range[0]:params.cin_range[1]].contiguous()
    if getattr(params, 'no_bias', False):
        b = torch.zeros_like(b)
    segs = getattr(params, 'cin_segs', None)
    if segs:                                   # input tensor with zero-padded channel blocks: spread the weight's columns
        assert sum(r for r, _ in segs) == w.shape[1], (segs, tuple(w.shape))
        w2 = torch.zeros(w.shape[0], sum(p for _, p in segs), w.shape[2], w.shape[3], dtype=w.dtype, device=w.device)
        rs = ps = 0
        for r, p in segs:
            w2[:, ps:ps + r] = w[:, rs:rs + r]
            rs, ps = rs + r, ps + p
        w = w2
    pad_out = int(getattr(params, 'pad_out', 0) or 0)
    if pad_out > w.shape[0]:                   # zero-padded output channels (zero rows, zero biases)
        w = torch.cat((w, torch.zeros(pad_out - w.shape[0], *w.shape[1:], dtype=w.dtype, device=w.device)), 0)
        b = torch.cat((b, torch.zeros(pad_out - b.shape[0], dtype=b.dtype, device=b.device)), 0)
    return w, b


def slab_of(cin, cout, groups):
    """(kslab, slab_mode) of a convolution -- see cpn_op_t::kslab."""
    if groups == 1:
        return cin, 0
    cg_in, cg_out = cin // groups, cout // groups
    assert cin == cout and cg_in == cg_out, 'grouped convolutions must have equal in/out widths'
    assert 64 % cg_in == 0 or cg_in % 64 == 0, 'group width must divide or be a multiple of 64'
    return max(64, cg_in), 1


def expand_grouped(w, groups):
    """[cout, cg, k, k] -> block-diagonal dense-slab weight [cout, kslab, k, k] (zeros off the diagonal blocks)."""
    cout, cg, kh, kw = w.shape
    cin = cg * groups
    kslab, _ = slab_of(cin, cout, groups)
    out = torch.zeros(cout, kslab, kh, kw, dtype=w.dtype, device=w.device)
    o = torch.arange(cout, device=w.device)
    grp = o // cg                                   # group of each output channel
    slab_base = (o // 64 * 64) // kslab * kslab     # first global input channel of the slab of o's 64-wide tile
    first = grp * cg - slab_base                    # local index of the group's first input channel
    for j in range(cg):
        out[o, first + j] = w[:, j]
    return out


def up2_weights(w, b):
    """Phase kernels of ``conv3x3(pad 1) o nearest-upsample(x2)`` (cpn_b200.h, CPN_CONV_UP2).  ``w`` [cout, cin, 3, 3],
    ``b`` [cout] -> ([4 * cout, cin, 3, 3], [4 * cout]): output pixel (2y + a, 2x + b') reads the up-sampled rows
    2y + a + t, t in {-1, 0, 1}, i.e. the source rows y + floor((a + t) / 2); taps that land on the same source pixel are
    summed.  Zero padding of the up-sampled image coincides with zero padding of the source, so the identity holds at the
    borders too (exactly, up to the rounding of the summed weights)."""
    cout, cin = w.shape[:2]
    out = torch.zeros(4, cout, cin, 3, 3, dtype=w.dtype)
    for a in range(2):
        for bb in range(2):
            for r in range(3):
                for s_ in range(3):
                    dy, dx = (a + r - 1) // 2, (bb + s_ - 1) // 2          # floor division
                    out[2 * a + bb, :, :, dy + 1, dx + 1] += w[:, :, r, s_]
    return out.reshape(4 * cout, cin, 3, 3), b.repeat(4)


def bilinear2_phase_matrix(k, phase):
    """1-D composition of ``F.interpolate(scale_factor=2, mode='bilinear', align_corners=False)`` with a k-tap correlation:
    ``A[r, d]`` such that ``sum_r W[r] U[2i + phase + r - k//2] = sum_d (W @ A)[d] L[i + d - D]`` for every output whose taps
    stay inside the image, with ``U[2m] = .25 L[m-1] + .75 L[m]``, ``U[2m+1] = .75 L[m] + .25 L[m+1]`` and ``D`` the low-res
    reach (2 for k = 7).  Returns (A [k, 2D + 1] float64, D)."""
    pad = k // 2
    D = (pad + 2) // 2
    A = torch.zeros(k, 2 * D + 1, dtype=torch.float64)
    for r in range(k):
        q = phase + r - pad
        m, odd = q // 2, q % 2                     # floor division: U row 2 (i + m) + odd
        if odd:
            A[r, m + D] += .75
            A[r, m + 1 + D] += .25
        else:
            A[r, m - 1 + D] += .25
            A[r, m + D] += .75
    return A, D


def bilinear2_weights(w, b):
    """``conv_kxk(pad k//2)(interpolate(x, x2, bilinear))`` as ONE (2D+1) x (2D+1) convolution on the low-res map with
    4 * cout output channels, phase (a, b') major (models/cpn.py:274-279 + the ReadOut's first convolution): ``w`` [cout, cin,
    k, k], ``b`` [cout] -> ([4 * cout, cin, 2D+1, 2D+1], [4 * cout]).  Exact wherever no tap of the k x k window touches the
    first / last row or column of the up-sampled image or its zero padding, i.e. outside a border of k//2 + 1 output
    pixels; the caller recomputes that border."""
    cout, cin, k, _ = w.shape
    w64 = w.double()
    outs = []
    for a in range(2):
        Ay, D = bilinear2_phase_matrix(k, a)
        for bb in range(2):
            Ax, _ = bilinear2_phase_matrix(k, bb)
            outs.append(torch.einsum('ocrs,rd,se->ocde', w64, Ay, Ax))
    return torch.cat(outs, 0).to(w.dtype).contiguous(), b.repeat(4)


def engine_for(op: LOp, fast, cin):
    if not fast:
        return L.ENGINE_SIMT
    kslab, _ = slab_of(cin, op.dst.c, op.params.groups)
    if kslab % 64 == 0 and op.dst.c % 64 == 0 and op.stride in (1, 2):
        return L.ENGINE_TCGEN05
    return L.ENGINE_SIMT


def pack_f16f8(wp, src_exp=0):
    """Weights of the 2-pass engine (cpn_b200.h, CPN_DT_F16F8).  ``wp``: fp32 ``[taps, cout, kslab]``.  Returns
    (uint8 ``[taps, cout, 4 * kslab]``, acc_scale): along K first the e4m3 block -- per 32-channel chunk 32 bytes
    ``e4m3(W_hi * 2^a)`` (pairing with the activations' lo8 bytes) then 32 bytes ``e4m3(W_lo * 2^(a+10))`` (pairing with
    hi8) -- then ``fp16(S * W_hi)``, with ``W_hi = fp16(W)``, ``W_lo = W - W_hi`` and one power of two ``a`` per layer that
    puts ``max|W| * 2^a`` in [64, 128).  Both halves of the 8-bit product then carry the scale of the main pass,
    ``S = 2^(8 + src_exp + a)``, so a single fp32 accumulator holds ``S * (A_hi W_hi + A_lo W_hi + A_hi W_lo)`` and the
    epilogue multiplies by ``acc_scale = 1 / S``."""
    taps, cout, kslab = wp.shape
    assert kslab % 32 == 0
    amax = float(wp.abs().max())
    a = (6 - math.floor(math.log2(amax)) if amax > 0 else 0) - max(int(src_exp), 0)
    a = max(min(a, 40), -40)
    hi = wp.half()
    hif = hi.float()
    lo = wp - hif
    s_main = 2. ** (8 + src_exp + a)
    w16 = (hif * s_main).half()
    assert bool(torch.isfinite(w16.float()).all()), 'fp16 overflow in the scaled main-pass weights'
    h8 = (hif * 2. ** a).clamp_(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).view(torch.uint8)
    l8 = (lo * 2. ** (a + 10)).clamp_(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).view(torch.uint8)
    w8 = torch.stack((h8.reshape(taps, cout, kslab // 32, 32), l8.reshape(taps, cout, kslab // 32, 32)), 3)
    out = torch.cat((w8.reshape(taps, cout, 2 * kslab), w16.view(torch.uint8).reshape(taps, cout, 2 * kslab)), 2)
    return out.contiguous(), 1. / s_main


class WeightPack:
    """Device blob with every conv / projection weight of a model in one precision mode.  ``split``: SPLIT_X3 -- weights
    of the 3-pass engine, ``(W_hi | W_hi | W_lo)`` along K with ``W_hi = fp16(W)``, ``W_lo = fp16(W - W_hi)``; SPLIT_F8 --
    weights of the 2-pass engine (``pack_f16f8``).  Packed on the CPU, uploaded with one copy."""

    def __init__(self, g: Tracer, sd, fast, device, split=SPLIT_NONE):
        split = int(split)
        chunks, off = [], 0
        self.entries = {}
        self.acc_scale = {}
        sd = {k: v.detach().to('cpu') for k, v in sd.items()}
        for i, op in enumerate(g.ops):
            if op.kind not in ('conv', 'proj'):
                continue
            w, b = fold_conv(sd, op.params)
            if op.kind == 'conv' and getattr(op, 'up2', False):
                w, b = up2_weights(w, b)
            if op.kind == 'conv' and getattr(op, 'bilin2', None):     # phase convolution of a bilinearly up-sampled input
                w, b = bilinear2_weights(w, b)
            if op.kind == 'conv' and getattr(op, 'gather', None):     # k x k conv as a 1x1 over gathered patches:
                k_, cin_ = op.gather                                   # K order [cin/64 blocks][k*k taps][64 channels]
                cout_ = w.shape[0]
                w = w.reshape(cout_, cin_ // 64, 64, k_, k_).permute(0, 1, 3, 4, 2).reshape(cout_, -1, 1, 1).contiguous()
            if op.kind == 'conv' and op.im2col is not None:   # [cout, cin, k, k] -> [cout, (r*k+s)*cin + c] padded
                if getattr(op, 'embed1x1', False):             # 1x1 on the raw input: the centre tap of the k x k matrix
                    k_ = op.im2col[0]
                    w1 = w
                    w = torch.zeros(w1.shape[0], w1.shape[1], k_, k_, dtype=w1.dtype)
                    w[:, :, k_ // 2, k_ // 2] = w1[:, :, 0, 0]
                cout_ = w.shape[0]
                wk = w.permute(0, 2, 3, 1).reshape(cout_, -1)
                w = torch.zeros(cout_, op.src.c, 1, 1, dtype=w.dtype)
                w[:, :wk.shape[1], 0, 0] = wk
            if op.kind == 'proj':
                wp = w.reshape(w.shape[0], w.shape[1]).contiguous().float()
                eng = -1
            else:
                cin = op.src.c
                eng = engine_for(op, fast, cin)
                if split and eng != L.ENGINE_TCGEN05:
                    raise NotImplementedError(f'{op.name}: the fp16x3 engine needs tensor-core-eligible layers')
                if op.params.groups > 1:
                    w = expand_grouped(w, op.params.groups)
                cout, kslab, kh, kw = w.shape
                if eng == L.ENGINE_TCGEN05:   # [R*S][cout][kslab] fp16
                    wp = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, kslab).contiguous()
                    if split == SPLIT_X3:
                        hi = wp.half()
                        lo = (wp - hi.float()).half()
                        wp = torch.cat((hi, hi, lo), 2).contiguous()
                    elif split == SPLIT_F8:
                        wp, self.acc_scale[i] = pack_f16f8(wp)
                    else:
                        wp = wp.half()
                else:                          # [R*S][kslab][cout] fp32
                    wp = w.permute(2, 3, 1, 0).reshape(kh * kw, kslab, cout).contiguous().float()
            wb = wp.view(torch.uint8).reshape(-1)
            bb = b.contiguous().float().view(torch.uint8).reshape(-1)
            w_off = off
            off = _align(off + wb.numel())
            b_off = off
            off = _align(off + bb.numel())
            chunks.append((w_off, wb))
            chunks.append((b_off, bb))
            self.entries[i] = (w_off, b_off, eng)
        host = torch.zeros(max(off, ALIGN), dtype=torch.uint8)
        for o, c in chunks:
            host[o:o + c.numel()] = c
        self.blob = host.to(device)          # the one host-to-device copy of the model's weights
        self.bytes = off


def _pitch(c, split=SPLIT_NONE):
    """Physical pixel pitch (elements) of a root buffer with c logical channels; split buffers hold (hi | lo) halves,
    fp16+e4m3 buffers (hi | 8-bit chunks of 32 channels)."""
    p = c if c % 4 == 0 else _align(c, 4)
    if split == SPLIT_X3:
        p = 2 * (p if p % 8 == 0 else _align(p, 8))
    elif split == SPLIT_F8:
        p = 2 * _align(p, 32)
    return p


def assign_arena(g: Tracer, elem_size, split=SPLIT_NONE):
    """First-fit placement of root buffers by live range.  Returns ({tensor id: byte offset of its root}, bytes)."""
    roots = []
    for t in g.tensors:
        if t.parent is None and not t.f32 and t.last >= 0:
            pitch = _pitch(t.c, split)
            roots.append((t.first, t.last, _align(g.n * t.h * t.w * pitch * elem_size), t))
    roots.sort(key=lambda r: (r[0], -r[2]))
    placed = []   # (offset, size, last)
    offsets = {}
    total = 0
    for first, last, size, t in roots:
        live = sorted((o, s) for o, s, l in placed if l >= first)
        pos = 0
        for o, s in live:
            if pos + size <= o:
                break
            pos = max(pos, o + s)
        placed.append((pos, size, last))
        offsets[t.id] = pos
        total = max(total, pos + size)
    return offsets, total


def _view(t: TT, n, root_off, elem_size, dtype):
    r, coff = t.root()
    split = {L.DT_F16X2: SPLIT_X3, L.DT_F16F8: SPLIT_F8}.get(dtype, SPLIT_NONE)
    assert split != SPLIT_F8 or coff % 32 == 0, 'fp16+e4m3 channel slices must start at multiples of 32'
    pitch = _pitch(r.c, split)
    v = L.View()
    v.offset = root_off.get(r.id, 0) + coff * elem_size
    v.n, v.h, v.w, v.c, v.pitch, v.dtype = n, t.h, t.w, t.c, pitch, dtype
    v.lo_delta = pitch // 2 if split else 0        # (hi block | lo block) inside every pixel of the root buffer
    return v


class Plan:
    """A compiled (architecture, N, H, W, precision) instance: C plan + arena + output buffers."""

    def __init__(self, g: Tracer, pack: WeightPack, fast, device, split=SPLIT_NONE):
        lib = L.load()
        split = int(split)
        self.g, self.pack, self.fast, self.device, self.split = g, pack, fast, device, split
        act_dt, es = ({SPLIT_NONE: L.DT_F16, SPLIT_X3: L.DT_F16X2, SPLIT_F8: L.DT_F16F8}[split], 2) if fast \
            else (L.DT_F32, 4)
        offsets, arena_bytes = assign_arena(g, es, split)
        self.arena = torch.empty(max(arena_bytes, ALIGN), dtype=torch.uint8, device=device)
        self.flags = torch.zeros(4, dtype=torch.int32, device=device)
        ops = (L.Op * len(g.ops))()
        kind_map = dict(prep=L.OP_PREP, conv=L.OP_CONV, maxpool=L.OP_MAXPOOL, upsample=L.OP_UPSAMPLE,
                        bilinear=L.OP_BILINEAR, proj=L.OP_PROJ)
        act_map = dict(none=L.ACT_NONE, relu=L.ACT_RELU, scaled_tanh=L.ACT_SCALED_TANH, sigmoid=L.ACT_SIGMOID)
        self.engines = []
        for i, lop in enumerate(g.ops):
            o = ops[i]
            o.kind = kind_map[lop.kind]
            o.out_binding = -1
            o.w_offset, o.b_offset = 0, -1
            if lop.dst.f32:
                r, coff = lop.dst.root()
                v = L.View()
                v.offset, v.n, v.h, v.w, v.c, v.pitch, v.dtype = coff * 4, g.n, lop.dst.h, lop.dst.w, lop.dst.c, r.c, L.DT_F32
                o.dst = v
                o.out_binding = lop.dst.binding
            else:
                o.dst = _view(lop.dst, g.n, offsets, es, act_dt)
            if lop.kind == 'prep' and lop.src is not None:   # im2col prep: logical input dims
                v = L.View()
                v.offset, v.n, v.h, v.w, v.c, v.pitch, v.dtype = 0, g.n, lop.src.h, lop.src.w, lop.src.c, lop.src.c, L.DT_F32
                o.src = v
            elif lop.src is not None:
                o.src = _view(lop.src, g.n, offsets, es, act_dt)
            else:
                o.src = o.dst
            if lop.res is not None:
                o.res = _view(lop.res, g.n, offsets, es, act_dt)
            o.r = o.s = (lop.k if (lop.kind != 'prep' or lop.im2col) else 0)
            o.stride, o.pad = lop.stride, lop.pad
            o.act, o.act_scale = act_map[lop.act], lop.act_scale
            if lop.kind in ('conv', 'proj'):
                w_off, b_off, eng = pack.entries[i]
                o.w_offset, o.b_offset = w_off, b_off
                if lop.kind == 'conv':
                    o.engine = eng
                    o.flags = L.CONV_UP2 if getattr(lop, 'up2', False) else 0
                    o.acc_scale = pack.acc_scale.get(i, 1.)
                    o.kslab, o.slab_mode = slab_of(lop.src.c, lop.dst.c, lop.params.groups)
                    self.engines.append(eng)
                else:
                    o.proj_cin_off, o.proj_cin = lop.cin_off, lop.cin
        # ---- fuse ReadOut projections into the epilogue of the tcgen05 convolution that feeds them ----
        self.fused = set()
        if fast:
            for i, lop in enumerate(g.ops):
                if lop.kind != 'conv' or ops[i].engine != L.ENGINE_TCGEN05 or lop.res is not None:
                    continue
                c = lop.dst.c
                bn = 256 if c % 256 == 0 else (128 if c % 128 == 0 else 64)
                nt = c // bn
                nxt = g.ops[i + 1:i + 1 + nt]
                if nt > 4 or len(nxt) != nt or lop.params.groups != 1:
                    continue
                ok = all(q.kind == 'proj' and q.src is lop.dst and q.cin == bn and q.cin_off == h * bn and q.dst.c <= 24
                         for h, q in enumerate(nxt))
                users = sum(1 for q in g.ops if q.src is lop.dst or q.res is lop.dst)
                if ok and users == nt and sum(q.dst.c for q in nxt) * bn * 4 <= 32768:
                    ops[i].fuse_next = nt
                    self.fused.update(range(i + 1, i + 1 + nt))
        self.ops = ops
        self.views = dict(offsets=offsets, es=es, act_dt=act_dt)
        handle = ctypes.c_void_p()
        L.check(lib.cpn_plan_create(ops, len(g.ops), L.ptr(pack.blob), pack.blob.numel(), L.ptr(self.arena),
                                    self.arena.numel(), L.ptr(self.flags), ctypes.byref(handle)), 'plan_create')
        self.handle = handle
        self.n_launches = lib.cpn_plan_num_launches(handle)
        hh, hw = g.head_hw
        self.out_shapes = OrderedDict()      # key -> shape; the list handed to the C plan is indexed by output binding
        for key, t in g.outputs.items():
            if key == 'scores':              # score maps keep the historical [N,h,w] shape for a single channel
                self.out_shapes[key] = (g.n, hh, hw) if t.c == 1 else (g.n, hh, hw, t.c)
            elif key == 'refinement':
                self.out_shapes[key] = (g.n,) + tuple(getattr(g, 'ref_hw', (g.h, g.w))) + (t.c,)
            else:
                self.out_shapes[key] = (g.n, t.h, t.w, t.c)
        self.out_bindings = {key: t.binding for key, t in g.outputs.items()}

    def new_outputs(self):
        """Output tensors indexed by binding (0 scores, 1 locfou, 2 refinement, 3 uncertainty; ``None`` where the plan has
        no such output, e.g. 'locfou' of a sparse-heads plan)."""
        outs = [None] * (max(self.out_bindings.values()) + 1)
        for key, shape in self.out_shapes.items():
            outs[self.out_bindings[key]] = torch.empty(shape, dtype=torch.float32, device=self.device)
        return outs

    def view_of(self, t: TT):
        """C view of logical tensor `t` inside this plan's arena (for entry points that read / fill plan tensors)."""
        return _view(t, self.g.n, self.views['offsets'], self.views['es'], self.views['act_dt'])

    def forward(self, x, input_format, outputs=None, first=0, end=None):
        """Enqueue the backbone + heads (or only ops [first, end)) on the current stream.  Returns the output list
        indexed by binding: [scores, locfou, refinement(, uncertainty)] (fp32)."""
        lib = L.load()
        outputs = self.new_outputs() if outputs is None else outputs
        arr = (ctypes.c_void_p * len(outputs))(*[None if o is None else o.data_ptr() for o in outputs])
        if first == 0 and end is None:
            L.check(lib.cpn_plan_forward(self.handle, L.ptr(x), input_format, arr, len(outputs), L.stream_ptr()),
                    'plan_forward')
        else:
            L.check(lib.cpn_plan_forward_range(self.handle, int(first), len(self.g.ops) if end is None else int(end),
                                               L.ptr(x), input_format, arr, len(outputs), L.stream_ptr()),
                    'plan_forward_range')
        return outputs

    def set_active_rows(self, rows):
        """Sparse-heads plans: compute only the 128-row blocks holding the first ``rows`` rows (< 0: all)."""
        L.check(L.load().cpn_plan_set_active_rows(self.handle, int(rows)), 'plan_set_active_rows')

    def forward_graph(self, x, input_format):
        """CUDA-graph replay of the whole plan (one graph launch instead of ``n_launches`` kernel launches): the input is
        copied into a static buffer, the graph is replayed on the current stream and the STATIC output tensors are
        returned -- they are overwritten by the next replay, so the caller consumes them first (the model's post-head
        chain does, on the same stream).  Captured on first use per (format, dtype, shape)."""
        key = (int(input_format), x.dtype, tuple(x.shape))
        graphs = self.__dict__.setdefault('_graphs', {})
        ent = graphs.get(key)
        if ent is None:
            static_x = torch.empty_like(x)
            static_out = self.new_outputs()
            static_x.copy_(x)
            self.forward(static_x, input_format, static_out)      # warm-up outside the capture (function attributes, ...)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.forward(static_x, input_format, static_out)
            ent = graphs[key] = (graph, static_x, static_out)
        graph, static_x, static_out = ent
        static_x.copy_(x, non_blocking=True)
        graph.replay()
        return static_out

    def run_op(self, index, x, input_format, outputs):
        lib = L.load()
        arr = (ctypes.c_void_p * len(outputs))(*[None if o is None else o.data_ptr() for o in outputs])
        L.check(lib.cpn_plan_run_op(self.handle, index, L.ptr(x), input_format, arr, len(outputs), L.stream_ptr()),
                'plan_run_op')

    def read_tensor(self, t: TT):
        """Debug: copy logical tensor `t` out of the arena as an NCHW fp32 tensor (valid right after forward only
        for tensors whose buffer has not been reused)."""
        es, dt = (2, torch.float16) if self.fast else (4, torch.float32)
        offsets, _ = assign_arena(self.g, es, self.split)
        r, coff = t.root()
        pitch = _pitch(r.c, self.split)
        nbytes = self.g.n * t.h * t.w * pitch * es
        base = offsets[r.id]
        buf = self.arena[base:base + nbytes].view(dt).reshape(self.g.n, t.h, t.w, pitch)
        out = buf[..., coff:coff + t.c].float()
        if self.split == SPLIT_X3:
            out = out + buf[..., pitch // 2 + coff:pitch // 2 + coff + t.c].float()
        elif self.split == SPLIT_F8:     # 8-bit block: per 32-channel chunk 32 lo8 bytes then 32 hi8 bytes
            b8 = buf[..., pitch // 2:].contiguous().view(torch.uint8).reshape(self.g.n, t.h, t.w, -1, 2, 32)
            lo8 = b8[..., 0, :].reshape(self.g.n, t.h, t.w, -1)[..., coff:coff + t.c]
            out = out + lo8.contiguous().view(torch.float8_e4m3fn).float() * 2. ** -8
        return out.permute(0, 3, 1, 2).contiguous()

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                L.load().cpn_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
