"""Model descriptions of the Contour Proposal Networks (U22 and the ResNet-family U-Net / FPN variants) as flat layer graphs.

The reference expresses these networks as ``nn.Module`` trees (all paths relative to /root/reference/celldetection):
``CpnU22`` (models/cpn.py:772, unet.py:405 ``U22``), ``CpnResNet18FPN`` (cpn.py:1250, fpn.py:240) and
``CpnResNeXt101UNet`` (cpn.py:930, unet.py:670).  Here one *tracer* walks the same structure and emits

* the ordered parameter/buffer specification with the reference's ``state_dict`` key names, and
* a flat list of layer ops on NHWC tensors (conv + folded BN + residual + ReLU, max-pool, nearest/bilinear resize, head
  projection) that ``plan.py`` lowers to the C ABI's ``cpn_op_t`` array.

Exact algebraic rewrites applied while tracing (each keeps per-pixel arithmetic identical):

* ``torch.cat((lateral, top_down), 1)`` (unet.py:219-224) is a shared NHWC buffer whose channel slices are written by
  the producers directly;
* the decoder's 1x1 "inner" convolution is applied *before* the nearest up-sampling it follows in the reference
  (unet.py:213-218): a 1x1 convolution commutes with nearest-neighbour replication, at a quarter of the work;
* torchvision's FPN ``lateral + interpolate(top_down)`` is the residual input of the lateral 1x1 convolution, read
  through the nearest index map;
* the three ReadOut heads that share the feature map (score / location / fourier, cpn.py:253-263) run as one
  convolution with concatenated output channels; Dropout2d is the identity in eval mode;
* FPN outputs CPN never reads (levels 2-4 and 'pool', fpn.py:67-76) are not computed, nor is the decoder's unused
  same-size ``'out'`` resample (unet.py:237).
"""
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

# ResNet-family encoders of models/resnet.py:330-487: name -> (layers, bottleneck, groups, width_per_group)
RESNETS = {
    'ResNet18': ((2, 2, 2, 2), False, 1, 64), 'ResNet34': ((3, 4, 6, 3), False, 1, 64),
    'ResNet50': ((3, 4, 6, 3), True, 1, 64), 'ResNet101': ((3, 4, 23, 3), True, 1, 64),
    'ResNet152': ((3, 8, 36, 3), True, 1, 64),
    'ResNeXt50': ((3, 4, 6, 3), True, 32, 4), 'ResNeXt101': ((3, 4, 23, 3), True, 32, 8),
    'ResNeXt152': ((3, 8, 36, 3), True, 32, 8),
    'WideResNet50': ((3, 4, 6, 3), True, 1, 128), 'WideResNet101': ((3, 4, 23, 3), True, 1, 128),
}
# Cpn<Encoder><Decoder> classes of models/cpn.py:930-1637 (U-Net: unet.py:591-716, FPN: fpn.py:240-322); the reference has
# no CpnWideResNet*UNet.  The three BASELINE configurations come first.
ARCHS = ('CpnU22', 'CpnResNet18FPN', 'CpnResNeXt101UNet') + tuple(
    f'Cpn{e}{d}' for d in ('UNet', 'FPN') for e in RESNETS
    if not (d == 'UNet' and e.startswith('Wide')) and f'Cpn{e}{d}' not in ('CpnResNet18FPN', 'CpnResNeXt101UNet')) + (
    'CpnWideU22',       # models/cpn.py:890-929, unet.py:497-524: U22 with doubled widths (128 ... 2048)
    'CpnResUNet',       # models/cpn.py:811-849, unet.py:434-464: U-Net whose encoder AND decoder blocks are ResBlocks
    'CpnSlimU22')       # models/cpn.py:851-889, unet.py:467-494: U22 with halved widths (32 ... 512); its 32-channel layers run
                        # zero-padded to the 64-wide tile (padded_conv)
# U-Net encoders of models/unet.py:405-524 by base width
U22_BASE = {'U22': 64, 'WideU22': 128, 'SlimU22': 32}


def split_arch(arch):
    """'CpnResNet50FPN' -> ('ResNet50', 'FPN'); 'CpnU22' -> ('U22', 'UNet')."""
    if arch in ('CpnU22', 'CpnWideU22', 'CpnResUNet', 'CpnSlimU22'):
        return arch[3:], 'UNet'
    for d in ('UNet', 'FPN'):
        if arch.endswith(d) and arch[3:-len(d)] in RESNETS:
            return arch[3:-len(d)], d
    raise ValueError(arch)


@dataclass
class TT:
    """Logical NHWC tensor of a trace."""
    id: int
    c: int
    h: int
    w: int
    parent: Optional['TT'] = None   # channel slice of `parent` at `c_off`
    c_off: int = 0
    f32: bool = False               # fp32 head output (bound to a caller buffer)
    binding: int = -1
    first: int = 1 << 30            # op index of first write
    last: int = -1                  # op index of last access
    segs: Optional[list] = None     # [(real channels, padded channels), ...] when the tensor carries zero-padded channel
                                    # blocks (widths that are not multiples of 64, CpnSlimU22); None = no padding

    def root(self):
        t, off = self, 0
        while t.parent is not None:
            off += t.c_off
            t = t.parent
        return t, off


@dataclass
class ConvParams:
    """Where a convolution's parameters live in the reference state_dict (None entries are absent)."""
    weight: List[str]                 # one or more keys; several = concatenated along the output-channel axis
    bias: List[Optional[str]]
    bn: List[Optional[str]]           # BatchNorm prefix per weight key
    groups: int = 1
    cin_range: Optional[Tuple[int, int]] = None   # contract over this slice of the (folded) weight's input channels only
    no_bias: bool = False                         # the folded bias is applied by another op of the same convolution
    pad_out: int = 0                              # > 0: zero rows / biases are appended up to this many output channels
    cin_segs: Optional[list] = None               # [(real, padded), ...]: the input tensor's channel blocks; the weight's input
                                                  # channels are spread accordingly (zero columns at the padded positions)


def channel_segs(t):
    """[(real, padded), ...] channel blocks of tensor ``t`` (one unpadded block unless ``t.segs`` says otherwise)."""
    return list(t.segs) if t.segs else [(t.c, t.c)]


def pad64(c):
    return (c + 63) // 64 * 64


def padded_conv(g, x, cout, k, params, **kw):
    """``g.conv`` for layers whose real widths are not multiples of 64 (CpnSlimU22's 32-channel layers): the output gets
    zero-padded channels up to the next multiple of 64 (zero weights and biases: exactly 0 after any activation used here) and
    the weight's input channels are spread over the padded blocks of ``x``.  Widths that are multiples of 64 pass through
    unchanged (same params object: the packed weights of every other architecture are bit-identical)."""
    from dataclasses import replace
    cp = pad64(cout)
    segs = channel_segs(x)
    spread = any(r != p for r, p in segs)
    if cp != cout or spread:
        params = replace(params, pad_out=cp if cp != cout else 0, cin_segs=segs if spread else None)
    out = g.conv(x, cp, k, params=params, **kw)
    if cp != cout:
        out.segs = [(cout, cp)]
    return out


@dataclass
class LOp:
    kind: str
    src: Optional[TT] = None
    dst: Optional[TT] = None
    res: Optional[TT] = None
    k: int = 1
    stride: int = 1
    pad: int = 0
    act: str = 'none'
    act_scale: float = 1.
    params: Optional[ConvParams] = None
    cin_off: int = 0     # proj: input channel slice
    cin: int = 0
    name: str = ''
    im2col: Optional[Tuple[int, int]] = None   # (k, cin) of the original convolution when it runs as a 1x1 on an
                                               # im2col'd input (tensor-core stem), or of the prep op producing it
    up2: bool = False                          # 3x3 conv on the 2x nearest-up-sampled src, computed from the low-res src
    gather: Optional[Tuple[int, int]] = None   # (k, cin): 1x1 conv over rows written by cpn_gather_patches, i.e. the
                                               # k x k convolution evaluated at selected pixels only
    bilin2: Optional[int] = None               # k: this (k//2 + 2 | 1)-tap convolution on the low-res src with 4 x the output
                                               # channels IS conv_kxk(interpolate(src, x2, bilinear)), phase-major columns
                                               # (plan.bilinear2_weights); exact outside an output border of k//2 + 1 pixels
    embed1x1: bool = False                     # a 1x1 convolution of the raw input that reads the SAME im2col matrix as the
                                               # k x k stem convolution: its weights sit at the centre tap's K positions


class Tracer:
    def __init__(self, n, h, w, stem_im2col=False, fuse_up2=False):
        self.n, self.h, self.w = n, h, w
        self.stem_im2col = stem_im2col   # run C_in<=4 stem convolutions on the tensor cores via an im2col'd input
        self.fuse_up2 = fuse_up2         # nearest x2 + 3x3 conv (the U-Net bridge block) as one phase-decomposed convolution
        self.tensors: List[TT] = []
        self.ops: List[LOp] = []
        self.spec = OrderedDict()   # state_dict key -> (shape, role)

    # ---- parameter specification ----------------------------------------------------------------------------------
    def conv_spec(self, key, cin, cout, k, bias, groups=1):
        self.spec[key + '.weight'] = ((cout, cin // groups, k, k), 'conv_w')
        if bias:
            self.spec[key + '.bias'] = ((cout,), 'conv_b')

    def bn_spec(self, key, c, residual_branch=False):
        self.spec[key + '.weight'] = ((c,), 'bn_w_res' if residual_branch else 'bn_w')
        self.spec[key + '.bias'] = ((c,), 'bn_b')
        self.spec[key + '.running_mean'] = ((c,), 'bn_rm')
        self.spec[key + '.running_var'] = ((c,), 'bn_rv')
        self.spec[key + '.num_batches_tracked'] = ((), 'bn_nbt')

    # ---- tensors / ops -------------------------------------------------------------------------------------------
    def tensor(self, c, h, w, **kw):
        t = TT(len(self.tensors), c, h, w, **kw)
        self.tensors.append(t)
        return t

    def _emit(self, op):
        i = len(self.ops)
        self.ops.append(op)
        for t in (op.src, op.res):
            if t is not None:
                t.last = max(t.last, i)
        op.dst.first = min(op.dst.first, i)
        op.dst.last = max(op.dst.last, i)
        return op.dst

    def external(self, c, h, w):
        """A tensor filled from outside the plan (no producing op): placed for the whole plan."""
        t = self.tensor(c, h, w)
        t.first, t.last = 0, 1 << 29
        return t

    def prep(self, c):
        t = self._emit(LOp('prep', dst=self.tensor(c, self.h, self.w), name='input'))
        t.is_input = True
        return t

    def stem_conv(self, x, cout, k, stride, act, params, name):
        """Convolution on the raw network input.  With ``stem_im2col`` the preceding prep op is turned into an im2col
        producer (K = k*k*c padded to 64) and the convolution becomes a 1x1 over it (same arithmetic, K reordered)."""
        pad = k // 2
        if not (self.stem_im2col and getattr(x, 'is_input', False) and self.ops[-1].kind == 'prep'):
            return self.conv(x, cout, k, stride=stride, act=act, params=params, name=name)
        prep = self.ops[-1]
        cin = x.c
        ho, wo = (x.h + 2 * pad - k) // stride + 1, (x.w + 2 * pad - k) // stride + 1
        kp = (k * k * cin + 63) // 64 * 64
        prep.src = TT(-1, cin, x.h, x.w)           # logical input (not an arena tensor)
        prep.k, prep.stride, prep.pad, prep.im2col = k, stride, pad, (k, cin)
        x.c, x.h, x.w = kp, ho, wo                 # the prep output is now the im2col matrix
        op = LOp('conv', src=x, dst=self.tensor(cout, ho, wo), k=1, stride=1, pad=0, act=act, params=params,
                 name=name, im2col=(k, cin))
        return self._emit(op)

    def stem_conv_1x1(self, x, cout, act, params, name):
        """A 1x1 convolution of the raw network input next to a k x k stem convolution (ResBlock's identity mapping,
        commons.py:292-296).  After ``stem_conv`` turned the input into an im2col matrix, the pixel itself is the centre tap's
        column block, so the 1x1 runs on the same matrix with its weights embedded there (zeros elsewhere)."""
        prep = next(o for o in self.ops if o.kind == 'prep')
        if not (self.stem_im2col and prep.im2col is not None and prep.dst is x):
            return self.conv(x, cout, 1, act=act, params=params, name=name)
        assert prep.stride == 1, 'embedded 1x1: the k x k stem convolution must have stride 1'
        op = LOp('conv', src=x, dst=self.tensor(cout, x.h, x.w), k=1, stride=1, pad=0, act=act, params=params, name=name,
                 im2col=prep.im2col, embed1x1=True)
        return self._emit(op)

    def conv(self, x, cout, k, stride=1, pad=None, act='none', res=None, params=None, name=''):
        pad = k // 2 if pad is None else pad
        ho, wo = (x.h + 2 * pad - k) // stride + 1, (x.w + 2 * pad - k) // stride + 1
        return self._emit(LOp('conv', src=x, dst=self.tensor(cout, ho, wo), res=res, k=k, stride=stride, pad=pad,
                              act=act, params=params, name=name))

    def conv_up2(self, x, cout, act='none', params=None, name=''):
        """``conv3x3(pad 1)(interpolate(x, scale_factor=2, mode='nearest'))`` without the up-sampled tensor
        (plan.up2_weights; tensor-core engines only)."""
        return self._emit(LOp('conv', src=x, dst=self.tensor(cout, 2 * x.h, 2 * x.w), k=3, stride=1, pad=1, act=act,
                              params=params, name=name, up2=True))

    def conv_cat_up2(self, lateral, top, cout, act, params, name):
        """``conv3x3(cat(lateral, interpolate(top, x2 nearest)))`` (models/unet.py:213-224 + the block's first conv) as two
        convolutions over the two halves of its input channels: the up-sampled half runs as phase kernels on the low-res
        `top` (each output phase needs 2 x 2 of the 3 x 3 low-res taps: 16 instead of 36 tap-products per source pixel) and
        writes the partial sum; the lateral half adds it through the residual input, with the bias and the activation.
        Neither the up-sampled tensor nor the concatenation exists."""
        from dataclasses import replace
        part = self.conv_up2(top, cout, act='none', name=name + '.up',
                             params=replace(params, cin_range=(lateral.c, lateral.c + top.c), no_bias=True))
        return self._emit(LOp('conv', src=lateral, dst=self.tensor(cout, lateral.h, lateral.w), res=part, k=3, stride=1,
                              pad=1, act=act, params=replace(params, cin_range=(0, lateral.c)), name=name))

    def maxpool(self, x, k, stride, pad):
        ho, wo = (x.h + 2 * pad - k) // stride + 1, (x.w + 2 * pad - k) // stride + 1
        return self._emit(LOp('maxpool', src=x, dst=self.tensor(x.c, ho, wo, segs=x.segs), k=k, stride=stride, pad=pad))

    def upsample(self, x, h, w):
        return self._emit(LOp('upsample', src=x, dst=self.tensor(x.c, h, w, segs=x.segs)))

    def bilinear(self, x, h, w):
        return self._emit(LOp('bilinear', src=x, dst=self.tensor(x.c, h, w, segs=x.segs)))

    def cat(self, a, b):
        """Concatenate along channels by making `a` and `b` slices of one buffer (both must be unplaced)."""
        assert a.parent is None and b.parent is None and (a.h, a.w) == (b.h, b.w)
        t = self.tensor(a.c + b.c, a.h, a.w)
        if a.segs or b.segs:
            t.segs = channel_segs(a) + channel_segs(b)
        a.parent, a.c_off = t, 0
        b.parent, b.c_off = t, a.c
        t.first = min(a.first, b.first)
        t.last = max(a.last, b.last)
        return t

    def proj(self, x, dst, cin_off, cin, params, act='none', act_scale=1., name=''):
        return self._emit(LOp('proj', src=x, dst=dst, cin_off=cin_off, cin=cin, params=params, act=act,
                              act_scale=act_scale, name=name))

    def move_ops(self, names, after_tensor):
        """Reorder: the ops named `names` (in their current order) move directly behind the LAST op that writes
        `after_tensor`; access ranges are recomputed.  Returns the index one past the moved block."""
        moved = [o for o in self.ops if o.name in names]
        rest = [o for o in self.ops if o.name not in names]
        pos = max(i for i, o in enumerate(rest) if o.dst is after_tensor) + 1
        self.ops = rest[:pos] + moved + rest[pos:]
        keep = {t.id: (t.first, t.last) for t in self.tensors if t.last >= (1 << 29) or (t.first == 0 and t.last >= (1 << 29))}
        for t in self.tensors:
            t.first, t.last = 1 << 30, -1
        for i, op in enumerate(self.ops):
            for t in (op.src, op.res):
                if t is not None:
                    t.last = max(t.last, i)
            op.dst.first = min(op.dst.first, i)
            op.dst.last = max(op.dst.last, i)
        for t in self.tensors:
            if t.id in keep:
                t.first, t.last = min(t.first, keep[t.id][0]), keep[t.id][1]
            if t.parent is not None:                       # cat: the shared buffer spans its slices
                r = t.root()[0]
                r.first, r.last = min(r.first, t.first), max(r.last, t.last)
        return pos + len(moved)

    def finalize(self):
        # propagate access ranges of slices to their roots
        for t in self.tensors:
            r, _ = t.root()
            if r is not t:
                r.first = min(r.first, t.first)
                r.last = max(r.last, t.last)


# ----------------------------------------------------------------------------------------------------------------------
# building blocks (state_dict layouts as dumped from the reference)
# ----------------------------------------------------------------------------------------------------------------------

def _conv_bn_act(g: Tracer, x, key, bn_key, cin, cout, k, stride=1, bias=True, groups=1, act='relu', res=None,
                 res_branch=False):
    g.conv_spec(key, cin, cout, k, bias, groups)
    if bn_key is not None:
        g.bn_spec(bn_key, cout, res_branch)
    p = ConvParams([key + '.weight'], [key + '.bias' if bias else None], [bn_key], groups)
    if getattr(x, 'is_input', False) and res is None and groups == 1:
        cp = pad64(cout)
        if cp != cout:
            from dataclasses import replace
            p = replace(p, pad_out=cp)
        out = g.stem_conv(x, cp, k, stride, act, p, key)
        if cp != cout:
            out.segs = [(cout, cp)]
        return out
    if groups == 1 and (cout % 64 or x.segs):
        return padded_conv(g, x, cout, k, p, stride=stride, act=act, res=res, name=key)
    return g.conv(x, cout, k, stride=stride, act=act, res=res, params=p, name=key)


def _two_conv_norm_relu(g, x, p, cin, cout, bias=True):
    """models/commons.py:120-149 (Sequential indices 0,1,(2),3,4,(5))"""
    x = _conv_bn_act(g, x, f'{p}.0', f'{p}.1', cin, cout, 3, bias=bias)
    return _conv_bn_act(g, x, f'{p}.3', f'{p}.4', cout, cout, 3, bias=bias)


def _res_block_spec(g, p, cin, cout):
    """Parameters of ``ResBlock`` (models/commons.py:259-359) in module order: downsample (ConvNorm 1x1, bias=False) when the
    widths differ, then block = conv3x3 (no bias), BN, act, conv3x3 (no bias), BN."""
    if cin != cout:
        g.conv_spec(f'{p}.downsample.0', cin, cout, 1, False)
        g.bn_spec(f'{p}.downsample.1', cout)
    g.conv_spec(f'{p}.block.0', cin, cout, 3, False)
    g.bn_spec(f'{p}.block.1', cout)
    g.conv_spec(f'{p}.block.3', cout, cout, 3, False)
    g.bn_spec(f'{p}.block.4', cout, True)


def _res_block_ops(g, x, p, cin, cout):
    """``act(block(x) + downsample(x))`` (commons.py:300-304): the second convolution adds the identity branch through its
    residual input.  On the raw network input the 3x3 becomes the im2col stem and the 1x1 identity mapping reads the same
    matrix (``Tracer.stem_conv_1x1``)."""
    pa = ConvParams([f'{p}.block.0.weight'], [None], [f'{p}.block.1'])
    pb = ConvParams([f'{p}.block.3.weight'], [None], [f'{p}.block.4'])
    pd = ConvParams([f'{p}.downsample.0.weight'], [None], [f'{p}.downsample.1'])
    if getattr(x, 'is_input', False):
        y = g.stem_conv(x, cout, 3, 1, 'relu', pa, f'{p}.block.0')
        idt = g.stem_conv_1x1(x, cout, 'none', pd, f'{p}.downsample.0')
    else:
        y = g.conv(x, cout, 3, act='relu', params=pa, name=f'{p}.block.0')
        idt = g.conv(x, cout, 1, act='none', params=pd, name=f'{p}.downsample.0') if cin != cout else x
    return g.conv(y, cout, 3, act='relu', res=idt, params=pb, name=f'{p}.block.3')


def _unet_encoder(g, x, p, cin, depth=5, base=64, block='two_conv'):
    """models/unet.py:29-58 (``block``: 'two_conv' = TwoConvNormRelu, 'res' = ResBlock)"""
    feats, chans = [], []
    for i in range(depth):
        cout = base * 2 ** i
        bp, bc = (f'{p}.0', cin) if i == 0 else (f'{p}.{i}.1', chans[-1])
        if i > 0:
            x = g.maxpool(x, 2, 2, 0)
        if block == 'res':
            _res_block_spec(g, bp, bc, cout)
            x = _res_block_ops(g, x, bp, bc, cout)
        else:
            x = _two_conv_norm_relu(g, x, bp, bc, cout)
        feats.append(x)
        chans.append(cout)
    return feats, chans


def _resnet_encoder(g, x, p, cin, kind):
    """models/resnet.py:265-290 with fused_initial=False; blocks :56-116; _make_layer :119-193."""
    layers, bottleneck, groups, width_per_group = RESNETS[kind]
    expansion = 4 if bottleneck else 1
    x = _conv_bn_act(g, x, f'{p}.0.0', f'{p}.0.1', cin, 64, 7, stride=2, bias=False)
    feats, chans = [x], [64]
    inplanes = 64
    for li, nblocks in enumerate(layers):
        planes = 64 * 2 ** li
        if li == 0:
            x = g.maxpool(x, 3, 2, 1)
        for bi in range(nblocks):
            bp = f'{p}.1.1.{bi}' if li == 0 else f'{p}.{li + 1}.{bi}'
            stride = 2 if (li > 0 and bi == 0) else 1
            outp = planes * expansion
            has_ds = stride != 1 or inplanes != outp
            if bottleneck:
                width = int(planes * (width_per_group / 64.)) * groups
                # register in module order: conv1,bn1,conv2,bn2,conv3,bn3,(downsample)
                y = _conv_bn_act(g, x, f'{bp}.conv1', f'{bp}.bn1', inplanes, width, 1, bias=False)
                y = _conv_bn_act(g, y, f'{bp}.conv2', f'{bp}.bn2', width, width, 3, stride=stride, bias=False,
                                 groups=groups)
                g.conv_spec(f'{bp}.conv3', width, outp, 1, False)
                g.bn_spec(f'{bp}.bn3', outp, True)
                idt = x
                if has_ds:
                    idt = _conv_bn_act(g, x, f'{bp}.downsample.0', f'{bp}.downsample.1', inplanes, outp, 1,
                                       stride=stride, bias=False, act='none')
                pr = ConvParams([f'{bp}.conv3.weight'], [None], [f'{bp}.bn3'])
                x = g.conv(y, outp, 1, act='relu', res=idt, params=pr, name=f'{bp}.conv3')
            else:
                y = _conv_bn_act(g, x, f'{bp}.conv1', f'{bp}.bn1', inplanes, planes, 3, stride=stride, bias=False)
                g.conv_spec(f'{bp}.conv2', planes, planes, 3, False)
                g.bn_spec(f'{bp}.bn2', planes, True)
                idt = x
                if has_ds:
                    idt = _conv_bn_act(g, x, f'{bp}.downsample.0', f'{bp}.downsample.1', inplanes, outp, 1,
                                       stride=stride, bias=False, act='none')
                pr = ConvParams([f'{bp}.conv2.weight'], [None], [f'{bp}.bn2'])
                x = g.conv(y, outp, 3, act='relu', res=idt, params=pr, name=f'{bp}.conv2')
            inplanes = outp
        feats.append(x)
        chans.append(inplanes)
    return feats, chans


def _unet_decoder(g, feats, chans, p, bridges, block='two_conv'):
    """models/unet.py:62-176 (bookkeeping) and :178-249 (forward), with the inner 1x1 conv commuted before the
    nearest up-sampling.  Returns the decoder outputs per level (index 0 = finest)."""
    in_list = [0] * bridges + list(chans)
    out_list = list(chans)                      # out_channels_list (unet.py:78-79, 100-104)
    nlev = len(in_list)
    # parameter registration order: inner_blocks (i = 1..nlev-1) interleaved with layer_blocks in the ctor loop; the
    # state_dict groups them by ModuleList, so register all inner blocks first, then all layer blocks.
    inner = {}
    for i in range(1, nlev):
        ouc = out_list[i - 1]
        inc = out_list[i] if i < nlev - 1 else in_list[i]
        if inc > 0 and ouc < inc:
            g.conv_spec(f'{p}.inner_blocks.{i - 1}', inc, ouc, 1, True)
            inner[i - 1] = (inc, ouc)
    blocks = {}
    for i in range(nlev - 1):
        lat = in_list[i]
        inc = min(out_list[i:i + 2])
        ouc = out_list[i]
        bias = lat > 0                           # bridge block = TwoConvNormRelu(bias=False) (unet.py:95-98)
        cin = inc + lat
        bp = f'{p}.layer_blocks.{i}'
        if block == 'res':                       # ResUNet: block_cls = ResBlock for the decoder too (unet.py:459-463)
            assert lat > 0, 'ResBlock decoders have no bridge levels'
            _res_block_spec(g, bp, cin, ouc)
            blocks[i] = (cin, ouc, bias)
            continue
        g.conv_spec(f'{bp}.0', cin, ouc, 3, bias)
        g.bn_spec(f'{bp}.1', ouc)
        g.conv_spec(f'{bp}.3', ouc, ouc, 3, bias)
        g.bn_spec(f'{bp}.4', ouc)
        blocks[i] = (cin, ouc, bias)
    depth = nlev - 1
    last = feats[-1]
    last_c = chans[-1]
    results = {}
    for i in range(depth - 1, -1, -1):
        lateral = feats[i - bridges] if i - bridges >= 0 else None
        top = last
        if i in inner:                           # forward applies inner_blocks[i] (unet.py:218)
            inc, ouc = inner[i]
            assert inc == last_c
            pr = ConvParams([f'{p}.inner_blocks.{i}.weight'], [f'{p}.inner_blocks.{i}.bias'], [None])
            top = padded_conv(g, top, ouc, 1, pr, act='none', name=f'{p}.inner_blocks.{i}')
            last_c = ouc
        cin, ouc, bias = blocks[i]
        bp = f'{p}.layer_blocks.{i}'
        if block == 'res':
            up = g.upsample(top, lateral.h, lateral.w)
            x = g.cat(lateral, up)               # cat_order 0: (lateral, top_down) (unet.py:219-224)
            assert x.c == cin, (x.c, cin)
            x = _res_block_ops(g, x, bp, cin, ouc)
            last, last_c = x, ouc
            results[i] = x
            continue
        pa = ConvParams([f'{bp}.0.weight'], [f'{bp}.0.bias' if bias else None], [f'{bp}.1'])
        pb = ConvParams([f'{bp}.3.weight'], [f'{bp}.3.bias' if bias else None], [f'{bp}.4'])
        if (lateral is not None and g.fuse_up2 and (lateral.h, lateral.w) == (2 * top.h, 2 * top.w) and
                lateral.parent is None and lateral.c % 64 == 0 and top.c % 64 == 0 and ouc % 256 == 0 and
                lateral.c + top.c == cin):
            x = g.conv_cat_up2(lateral, top, ouc, 'relu', pa, f'{bp}.0')
            x = g.conv(x, ouc, 3, act='relu', params=pb, name=f'{bp}.3')
            last, last_c = x, ouc
            results[i] = x
            continue
        if lateral is not None:
            up = g.upsample(top, lateral.h, lateral.w)
            x = g.cat(lateral, up)               # cat_order 0: (lateral, top_down) (unet.py:219-224)
        elif g.fuse_up2 and top.c % 64 == 0 and ouc % 64 == 0:
            # bridge level: nearest x2 followed by the block's first 3x3 conv -> four phase kernels on the low-res map
            assert top.c == cin, (top.c, cin)
            x = g.conv_up2(top, ouc, act='relu', params=pa, name=f'{bp}.0')
            x = g.conv(x, ouc, 3, act='relu', params=pb, name=f'{bp}.3')
            last, last_c = x, ouc
            results[i] = x
            continue
        else:
            x = g.upsample(top, top.h * 2, top.w * 2)
        assert sum(r for r, _ in channel_segs(x)) == cin, (channel_segs(x), cin)
        x = padded_conv(g, x, ouc, 3, pa, act='relu', name=f'{bp}.0')
        x = padded_conv(g, x, ouc, 3, pb, act='relu', name=f'{bp}.3')
        last, last_c = x, ouc
        results[i] = x
    return results, out_list


def _fpn_decoder(g, feats, chans, p, fpn_channels=256, needed=(0, 1)):
    """torchvision FeaturePyramidNetwork.forward with models/fpn.py:79-134 blocks (conv + bias, no norm/activation).
    Only the levels CPN reads are finished with their 3x3 output convolution."""
    n = len(feats)
    for i in range(n):
        g.conv_spec(f'{p}.inner_blocks.{i}.0', chans[i], fpn_channels, 1, True)
    for i in range(n):
        g.conv_spec(f'{p}.layer_blocks.{i}.0', fpn_channels, fpn_channels, 3, True)
    results = {}
    last = None
    for i in range(n - 1, -1, -1):
        pr = ConvParams([f'{p}.inner_blocks.{i}.0.weight'], [f'{p}.inner_blocks.{i}.0.bias'], [None])
        last = g.conv(feats[i], fpn_channels, 1, act='none', res=last, params=pr, name=f'{p}.inner_blocks.{i}')
        if i in needed:
            po = ConvParams([f'{p}.layer_blocks.{i}.0.weight'], [f'{p}.layer_blocks.{i}.0.bias'], [None])
            results[i] = g.conv(last, fpn_channels, 3, act='none', params=po, name=f'{p}.layer_blocks.{i}')
    return results


def _read_out_spec(g, p, cin, cmid, cout, k=7):
    """models/commons.py:461-511: block.0 conv kxk (bias), block.1 BN, block.2 act, block.3 dropout, block.4 conv 1x1."""
    g.conv_spec(f'{p}.block.0', cin, cmid, k, True)
    g.bn_spec(f'{p}.block.1', cmid)
    g.conv_spec(f'{p}.block.4', cmid, cout, 1, True)


HEAD_KERNEL_KEYS = ('score', 'location', 'fourier', 'uncertainty', 'refinement')


def trace(arch, n, h, w, in_channels=3, order=5, score_channels=1, refinement_margin=3., refinement_buckets=1,
          uncertainty_head=False, stem_im2col=False, kernel_sizes=None, contour_head_channels=None,
          refinement_head_channels=None, contour_head_stride=1, refinement_head_stride=1, refinement_full_res=True,
          fpn_channels=256, fuse_up2=False, sparse_heads=False, phase_refinement=False):
    """Trace architecture `arch` for an [n, in_channels, h, w] input.  Returns the Tracer; ``g.outputs`` maps
    'scores' / 'locfou' / 'refinement' (/ 'uncertainty') to fp32 output tensors (bindings 0 / 1 / 2 (/ 3)).

    Variants of models/cpn.py:177-234: ``score_channels`` > 1 widens the score head (classes > 2, :372),
    ``refinement_buckets`` > 1 widens the refinement head to 2 * buckets channels (:222-234), ``uncertainty_head``
    adds a fourth ReadOut with 4 sigmoid outputs on the head features (:208-219).  Shape-changing head options:
    ``kernel_sizes`` (dict over HEAD_KERNEL_KEYS, default 7 each: ``kernel_size_<head>``, :179-229), the ReadOut mid
    widths ``contour_head_channels`` / ``refinement_head_channels`` (default: the input width), the head strides
    ``contour_head_stride`` / ``refinement_head_stride``, ``refinement_full_res`` (:277-279) and ``fpn_channels``
    (fpn.py:240-322).  Contour heads that share a kernel size run as one merged convolution; the refinement tensor is
    produced at the head's own resolution ``g.ref_hw`` (the caller resizes it to the input size when they differ, :279).

    ``sparse_heads``: the location and fourier heads are NOT part of this plan -- they are only read at proposals
    (cpn.py:620-623) and are evaluated there by ``trace_sparse_heads`` after the selection; the head feature map stays
    alive until the end of the plan (``g.head_feat``) and ``g.outputs`` has no 'locfou'.  ``g.sparse`` tells whether the
    option could be applied (stride-1 contour heads, equal location / fourier kernel sizes).

    ``phase_refinement``: when the refinement features are up-sampled x2 to the input size before a 7x7 head (:274-279; the
    FPN models), the bilinear op and the 7x7 convolution at full resolution are replaced by ONE 5x5 convolution on the
    low-res features with 4 x the mid channels (the four output phases; 25 instead of 49 taps per output, no full-res
    feature tensor) and the fused projection per phase.  'refinement' is then the phase-packed record map [n, h/2, w/2,
    4 * 2B] (``g.ref_phase``); the caller shuffles it to full resolution and recomputes the 4-pixel image border, where the
    identity does not hold, with ``trace_ring_strip`` plans on cropped features (kept alive: ``g.ref_phase['feat']``)."""
    assert arch in ARCHS, arch
    assert score_channels >= 1 and refinement_buckets >= 1
    ks = dict.fromkeys(HEAD_KERNEL_KEYS, 7)
    ks.update(kernel_sizes or {})
    assert all(int(k) % 2 == 1 and 1 <= int(k) <= 15 for k in ks.values()), 'head kernel sizes must be odd, <= 15'
    g = Tracer(n, h, w, stem_im2col=stem_im2col, fuse_up2=fuse_up2)
    g.spec['order_weights'] = ((order, 1), 'order_weights')   # buffer of CPN (cpn.py:406-412)
    bb = 'core.backbone'
    x = g.prep(in_channels)
    enc, dec = split_arch(arch)
    if enc in U22_BASE or enc == 'ResUNet':
        block = 'res' if enc == 'ResUNet' else 'two_conv'
        feats, chans = _unet_encoder(g, x, f'{bb}.body', in_channels, base=U22_BASE.get(enc, 64), block=block)
        res, out_ch = _unet_decoder(g, feats, chans, f'{bb}.unet', bridges=0, block=block)
        head_feat, head_c, ref_feat, ref_c = res[1], out_ch[1], res[0], out_ch[0]
    elif dec == 'UNet':
        feats, chans = _resnet_encoder(g, x, f'{bb}.body', in_channels, enc)
        res, out_ch = _unet_decoder(g, feats, chans, f'{bb}.unet', bridges=1)
        head_feat, head_c, ref_feat, ref_c = res[1], out_ch[1], res[0], out_ch[0]
    else:
        feats, chans = _resnet_encoder(g, x, f'{bb}.body', in_channels, enc)
        res = _fpn_decoder(g, feats, chans, f'{bb}.fpn', fpn_channels=fpn_channels)
        head_feat, head_c, ref_feat, ref_c = res[1], fpn_channels, res[0], fpn_channels
    # ---- heads (models/cpn.py:177-234, 238-283); module order: score, location, fourier, (uncertainty), refinement ----
    heads = [('core.score_head', score_channels, 'none', int(ks['score'])),
             ('core.location_head', 2, 'none', int(ks['location'])),
             ('core.fourier_head', order * 4, 'none', int(ks['fourier']))]
    if uncertainty_head:
        heads.append(('core.uncertainty_head', 4, 'sigmoid', int(ks['uncertainty'])))
    head_mid = int(contour_head_channels) if contour_head_channels else head_c
    ref_mid = int(refinement_head_channels) if refinement_head_channels else ref_c
    for hp, co, _, k in heads:
        _read_out_spec(g, hp, head_c, head_mid, co, k)
    _read_out_spec(g, 'core.refinement_head', ref_c, ref_mid, 2 * refinement_buckets, int(ks['refinement']))
    cs = int(contour_head_stride)
    hh = (head_feat.h + 2 * (heads[0][3] // 2) - heads[0][3]) // cs + 1
    hw_ = (head_feat.w + 2 * (heads[0][3] // 2) - heads[0][3]) // cs + 1
    scores = g.tensor(score_channels, hh, hw_, f32=True, binding=0)
    locfou = g.tensor(2 + 4 * order, hh, hw_, f32=True, binding=1)
    loc_t = g.tensor(2, hh, hw_, f32=True, parent=locfou, c_off=0, binding=1)
    fou_t = g.tensor(4 * order, hh, hw_, f32=True, parent=locfou, c_off=2, binding=1)
    dsts = [scores, loc_t, fou_t]
    uncertainty = None
    if uncertainty_head:
        uncertainty = g.tensor(4, hh, hw_, f32=True, binding=3)
        dsts.append(uncertainty)
    g.sparse = bool(sparse_heads and cs == 1 and ks['location'] == ks['fourier'] and head_c % 64 == 0)
    g.head_feat, g.head_mid, g.head_k = head_feat, head_mid, int(ks['location'])
    dense = list(zip(heads, dsts))
    if g.sparse:
        dense = [(hd, dst) for hd, dst in dense if hd[0] not in ('core.location_head', 'core.fourier_head')]
        head_feat.last = 1 << 29           # read by cpn_gather_patches after the plan
    # heads with the same kernel size share one convolution (concatenated output channels, in module order)
    for k in sorted({hd[3] for hd, _ in dense}, reverse=True):
        grp = [(hd, dst) for hd, dst in dense if hd[3] == k]
        pm = ConvParams([f'{hd[0]}.block.0.weight' for hd, _ in grp], [f'{hd[0]}.block.0.bias' for hd, _ in grp],
                        [f'{hd[0]}.block.1' for hd, _ in grp])
        name = 'heads.block.0' if len(grp) == len(dense) else 'heads.block.0.k%d' % k
        mid = g.conv(head_feat, len(grp) * head_mid, k, stride=cs, act='relu', params=pm, name=name)
        assert (mid.h, mid.w) == (hh, hw_)
        for j, ((hp, co, act, _), dst) in enumerate(grp):
            pp = ConvParams([f'{hp}.block.4.weight'], [f'{hp}.block.4.bias'], [None])
            g.proj(mid, dst, j * head_mid, head_mid, pp, act=act, name=f'{hp}.block.4')
    c2 = 2 * refinement_buckets
    pr = ConvParams(['core.refinement_head.block.0.weight'], ['core.refinement_head.block.0.bias'],
                    ['core.refinement_head.block.1'])
    pp = ConvParams(['core.refinement_head.block.4.weight'], ['core.refinement_head.block.4.bias'], [None])
    g.ref_phase = None
    if (phase_refinement and refinement_full_res and (2 * ref_feat.h, 2 * ref_feat.w) == (h, w) and
            int(ks['refinement']) == 7 and int(refinement_head_stride) == 1 and ref_mid % 256 == 0 and ref_c % 64 == 0 and
            ref_feat.parent is None and ref_feat.h >= 8 and ref_feat.w >= 8):
        rmid = g._emit(LOp('conv', src=ref_feat, dst=g.tensor(4 * ref_mid, ref_feat.h, ref_feat.w), k=5, stride=1, pad=2,
                           act='relu', params=pr, name='core.refinement_head.block.0', bilin2=7))
        refinement = g.tensor(4 * c2, rmid.h, rmid.w, f32=True, binding=2)
        for ph in range(4):
            dst = g.tensor(c2, rmid.h, rmid.w, f32=True, parent=refinement, c_off=ph * c2, binding=2)
            g.proj(rmid, dst, ph * ref_mid, ref_mid, pp, act='scaled_tanh', act_scale=float(refinement_margin),
                   name=f'core.refinement_head.block.4.phase{ph}')
        ref_feat.last = 1 << 29            # cropped into the border-strip plans after the plan
        g.ref_phase = dict(feat=ref_feat, c=ref_c, mid=ref_mid, k=7, c2=c2, margin=float(refinement_margin))
    else:
        if refinement_full_res and (ref_feat.h, ref_feat.w) != (h, w):       # refinement_full_res (cpn.py:277-278)
            ref_feat = g.bilinear(ref_feat, h, w)
        rmid = padded_conv(g, ref_feat, ref_mid, int(ks['refinement']), pr, stride=int(refinement_head_stride), act='relu',
                           name='core.refinement_head.block.0')
        refinement = g.tensor(c2, rmid.h, rmid.w, f32=True, binding=2)
        if rmid.segs:                      # zero-padded mid channels: the projection reads them with zero weights
            from dataclasses import replace
            pp = replace(pp, cin_segs=rmid.segs)
        g.proj(rmid, refinement, 0, rmid.c, pp, act='scaled_tanh', act_scale=float(refinement_margin),
               name='core.refinement_head.block.4')
    g.outputs = OrderedDict(scores=scores, locfou=locfou, refinement=refinement)
    if g.sparse:
        del g.outputs['locfou']
    if uncertainty is not None:
        g.outputs['uncertainty'] = uncertainty
    g.head_hw = (hh, hw_)
    g.ref_hw = (rmid.h, rmid.w)
    g.split_index = None
    if g.sparse:
        # the dense heads (score, uncertainty) run as soon as their features exist: the host reads the proposal count
        # while the GPU is still busy with the full-resolution branch and the refinement head (ops [split_index, end))
        head_ops = [o.name for o in g.ops if o.name.startswith('heads.block.0') or
                    o.name in ('core.score_head.block.4', 'core.uncertainty_head.block.4')]
        g.split_index = g.move_ops(head_ops, head_feat)
    g.finalize()
    return g


def trace_ring_strip(n, hs, ws, ref_c, ref_mid, c2, margin, k=7):
    """The refinement ReadOut on a cropped strip of the low-res refinement features, by the plain path (bilinear x2, k x k
    convolution, projection; models/cpn.py:274-279): recomputes the image border of a ``phase_refinement`` plan.  External
    input [n, hs, ws, ref_c] -> 'refinement' [n, 2 hs, 2 ws, c2]."""
    g = Tracer(n, hs, ws)
    a = g.external(ref_c, hs, ws)
    u = g.bilinear(a, 2 * hs, 2 * ws)
    pr = ConvParams(['core.refinement_head.block.0.weight'], ['core.refinement_head.block.0.bias'],
                    ['core.refinement_head.block.1'])
    rmid = g.conv(u, ref_mid, k, act='relu', params=pr, name='core.refinement_head.block.0')
    out = g.tensor(c2, 2 * hs, 2 * ws, f32=True, binding=0)
    pp = ConvParams(['core.refinement_head.block.4.weight'], ['core.refinement_head.block.4.bias'], [None])
    g.proj(rmid, out, 0, ref_mid, pp, act='scaled_tanh', act_scale=float(margin), name='core.refinement_head.block.4')
    g.outputs = OrderedDict(refinement=out)
    g.input_tensor = a
    g.head_hw = g.ref_hw = (2 * hs, 2 * ws)
    g.finalize()
    return g


def trace_sparse_heads(rows, head_c, head_mid, k, order):
    """The location + fourier ReadOut heads (models/commons.py:461-511) on ``rows`` gathered k x k x head_c patches
    (``cpn_gather_patches``; rows is a multiple of 128): ONE 1x1 convolution over K = k*k*head_c (the dense k x k
    convolution's own contraction, in its K order) + BN + ReLU with the two final projections fused -> 'locfou' records
    [rows, 2 + 4*order].  The gathered matrix is an external tensor [1, rows/16, 16, k*k*head_c]."""
    assert rows % 128 == 0
    g = Tracer(1, rows // 16, 16)
    a = g.external(k * k * head_c, rows // 16, 16)
    heads = [('core.location_head', 2), ('core.fourier_head', order * 4)]
    pm = ConvParams([f'{hp}.block.0.weight' for hp, _ in heads], [f'{hp}.block.0.bias' for hp, _ in heads],
                    [f'{hp}.block.1' for hp, _ in heads])
    mid = g._emit(LOp('conv', src=a, dst=g.tensor(2 * head_mid, rows // 16, 16), k=1, stride=1, pad=0, act='relu',
                      params=pm, name='sparse_heads.block.0', gather=(k, head_c)))
    locfou = g.tensor(2 + 4 * order, rows // 16, 16, f32=True, binding=0)
    loc_t = g.tensor(2, rows // 16, 16, f32=True, parent=locfou, c_off=0, binding=0)
    fou_t = g.tensor(4 * order, rows // 16, 16, f32=True, parent=locfou, c_off=2, binding=0)
    for j, ((hp, co), dst) in enumerate(zip(heads, (loc_t, fou_t))):
        pp = ConvParams([f'{hp}.block.4.weight'], [f'{hp}.block.4.bias'], [None])
        g.proj(mid, dst, j * head_mid, head_mid, pp, name=f'{hp}.block.4')
    g.outputs = OrderedDict(locfou=locfou)
    g.input_tensor = a
    g.head_hw = (rows // 16, 16)
    g.finalize()
    return g


def conv_flops(g: Tracer):
    """Executed convolution FLOPs of a trace (2 * output elements * (C_in / groups) * k * k),
    including the head projections."""
    total = 0
    for op in g.ops:
        if op.kind == 'conv':
            kk = op.im2col[0] ** 2 * op.im2col[1] if op.im2col else (op.src.c // op.params.groups) * op.k * op.k
            if op.embed1x1:
                kk = op.im2col[1]
            if op.bilin2:                 # counted in the reference formulation: k x k taps at the full resolution
                kk = op.src.c * op.bilin2 ** 2
            if op.up2 and op.params.cin_range is not None:
                kk = kk * 4 // 9          # phase N tiles issue their 2 x 2 taps only (the reference formulation: 9 taps)
            if op.gather:
                kk = op.gather[0] ** 2 * op.gather[1]
            total += 2 * g.n * op.dst.h * op.dst.w * op.dst.c * kk
        elif op.kind == 'proj':
            total += 2 * g.n * op.dst.h * op.dst.w * op.dst.c * op.cin
    return total
