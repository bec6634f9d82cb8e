"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of celldetection's CPN inference hot path.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  ``celldetection_b200`` never does.

It restates, with plain ``torch.nn.functional`` CPU ops (fp32) and numpy, what the reference computes on the path
``cd.models.CPN.forward`` (``/root/reference/celldetection/models/cpn.py:561-734``) for the three configured
architectures, directly from a reference-format ``state_dict``:

=====================  ==========================================================================================
oracle function        reference code it follows (relative to /root/reference/celldetection)
=====================  ==========================================================================================
``normalize``          models/commons.py:694-700 (range assert + (x-mean)/std, mean 0 / std 1 here)
``two_conv_norm_relu`` models/commons.py:120-149
``unet_encoder``       models/unet.py:29-58 (U22: 5 levels, MaxPool2x2 from level 1)
``resnet_encoder``     models/resnet.py:265-290 (stem), :56-116 (+ torchvision BasicBlock/Bottleneck.forward), :119-193
``unet_decoder``       models/unet.py:178-249 (forward) with the ctor bookkeeping of :62-176
``fpn_decoder``        torchvision.ops.FeaturePyramidNetwork.forward, models/fpn.py:50-76 (LastLevelMaxPool), :79-134
``read_out``           models/commons.py:461-511
``cpn_core5``          models/cpn.py:238-283 (``cpn_core`` = without the uncertainty map)
``cpn_post``           models/cpn.py:575-734 (all inference variants), ops/cpn.py:15-165, :189-227, :238-255
``nms``                torch.ops.torchvision.nms semantics (third party, torchvision 0.26)
``fouriers2contours``  ops/cpn.py:44-95
``get_tiling_slices``  util/util.py:1305-1354
``remove_border_contours`` ops/cpn.py:258-290
``filter_contours_by_stitching_rule`` ops/cpn.py:293-325
``apply_model``        celldetection_scripts/cpn_inference.py:311-411 (one model, masks / point masks, 'nms' stitching)
``apply_models``       cpn_inference.py:354-429 (ensemble: per-model results, box voting, final NMS)
``filter_by_box_voting`` ops/boxes.py:53-83 (+ torchvision box_iou)
=====================  ==========================================================================================

Pinning: the reference ships no golden vectors for this path, so the oracle is pinned against
outputs of the reference itself, executed in the build container through ``oracle/ref_shim.py``; the generated
vectors live in ``tests/golden`` (script: ``oracle/make_golden.py``; check: ``tests/test_oracle_golden.py``).
"""
from collections import OrderedDict
import math

import numpy as np
import torch
import torch.nn.functional as F

NMS_BATCH_SIZE = 50000  # ops/cpn.py:12

# ResNet-family encoders (models/resnet.py:330-487): name -> (layers, bottleneck, groups, width_per_group)
RESNETS = {
    'ResNet18': ((2, 2, 2, 2), False, 1, 64), 'ResNet34': ((3, 4, 6, 3), False, 1, 64),
    'ResNet50': ((3, 4, 6, 3), True, 1, 64), 'ResNet101': ((3, 4, 23, 3), True, 1, 64),
    'ResNet152': ((3, 8, 36, 3), True, 1, 64),
    'ResNeXt50': ((3, 4, 6, 3), True, 32, 4), 'ResNeXt101': ((3, 4, 23, 3), True, 32, 8),
    'ResNeXt152': ((3, 8, 36, 3), True, 32, 8),
    'WideResNet50': ((3, 4, 6, 3), True, 1, 128), 'WideResNet101': ((3, 4, 23, 3), True, 1, 128),
}
ARCHS = {'CpnU22': dict(decoder='unet', encoder='unet', head_key='1', ref_key='0'),
         # models/cpn.py:890-929 / unet.py:497-524: U22 with doubled channel widths (the widths are read off the state_dict)
         'CpnWideU22': dict(decoder='unet', encoder='unet', head_key='1', ref_key='0'),
         # models/cpn.py:811-849 / unet.py:434-464: U-Net of ResBlocks (the block type is read off the state_dict keys)
         'CpnResUNet': dict(decoder='unet', encoder='unet', head_key='1', ref_key='0'),
         # models/cpn.py:851-889 / unet.py:467-494: U22 with halved channel widths
         'CpnSlimU22': dict(decoder='unet', encoder='unet', head_key='1', ref_key='0')}
for _e in RESNETS:                       # models/cpn.py:930-1637 (the reference has no CpnWideResNet*UNet)
    ARCHS[f'Cpn{_e}FPN'] = dict(decoder='fpn', encoder=_e, head_key='1', ref_key='0')
    if not _e.startswith('Wide'):
        ARCHS[f'Cpn{_e}UNet'] = dict(decoder='unet', encoder=_e, head_key='1', ref_key='0')


# ----------------------------------------------------------------------------------------------------------------------
# Building blocks
# ----------------------------------------------------------------------------------------------------------------------

def _conv(x, sd, key, stride=1, padding=0, groups=1):
    return F.conv2d(x, sd[key + '.weight'], sd.get(key + '.bias'), stride=stride, padding=padding, groups=groups)


def _bn(x, sd, key, eps=1e-5):
    return F.batch_norm(x, sd[key + '.running_mean'], sd[key + '.running_var'], sd[key + '.weight'],
                        sd[key + '.bias'], training=False, eps=eps)


def normalize(x):
    """models/commons.py:694-700 with mean 0, std 1, assert_range (0, 1)."""
    assert torch.all(x >= 0.) and torch.all(x <= 1.), 'Inputs should be in interval (0.0, 1.0)'
    return (x - 0.) / 1.


def two_conv_norm_relu(x, sd, p):
    """conv3x3, BN, ReLU, conv3x3, BN, ReLU with Sequential indices 0,1,3,4 (models/commons.py:120-149)."""
    x = F.relu(_bn(_conv(x, sd, f'{p}.0', padding=1), sd, f'{p}.1'))
    x = F.relu(_bn(_conv(x, sd, f'{p}.3', padding=1), sd, f'{p}.4'))
    return x


def res_block(x, sd, p):
    """models/commons.py:259-359 ``ResBlock``: act(block(x) + downsample(x)); downsample = ConvNorm 1x1 (bias=False) when the
    widths differ, block = conv3x3, BN, ReLU, conv3x3, BN (both convolutions without bias)."""
    idt = x
    if f'{p}.downsample.0.weight' in sd:
        idt = _bn(_conv(x, sd, f'{p}.downsample.0'), sd, f'{p}.downsample.1')
    out = F.relu(_bn(_conv(x, sd, f'{p}.block.0', padding=1), sd, f'{p}.block.1'))
    out = _bn(_conv(out, sd, f'{p}.block.3', padding=1), sd, f'{p}.block.4')
    return F.relu(out + idt)


def unet_block(x, sd, p):
    """The U-Net's ``block_cls`` at prefix ``p``: ResBlock (ResUNet, unet.py:455-463) or TwoConvNormRelu."""
    return res_block(x, sd, p) if f'{p}.block.0.weight' in sd else two_conv_norm_relu(x, sd, p)


def unet_encoder(x, sd, p, depth=5):
    """models/unet.py:29-58: level 0 is the bare block, levels > 0 are Sequential(MaxPool(2, 2), block)."""
    feats = OrderedDict()
    for i in range(depth):
        if i == 0:
            x = unet_block(x, sd, f'{p}.0')
        else:
            x = F.max_pool2d(x, 2, 2)
            x = unet_block(x, sd, f'{p}.{i}.1')
        feats[str(i)] = x
    return feats


def _basic_block(x, sd, p, stride):
    idt = x
    out = F.relu(_bn(_conv(x, sd, f'{p}.conv1', stride=stride, padding=1), sd, f'{p}.bn1'))
    out = _bn(_conv(out, sd, f'{p}.conv2', padding=1), sd, f'{p}.bn2')
    if f'{p}.downsample.0.weight' in sd:
        idt = _bn(_conv(x, sd, f'{p}.downsample.0', stride=stride), sd, f'{p}.downsample.1')
    return F.relu(out + idt)


def _bottleneck(x, sd, p, stride, groups):
    idt = x
    out = F.relu(_bn(_conv(x, sd, f'{p}.conv1'), sd, f'{p}.bn1'))
    out = F.relu(_bn(_conv(out, sd, f'{p}.conv2', stride=stride, padding=1, groups=groups), sd, f'{p}.bn2'))
    out = _bn(_conv(out, sd, f'{p}.conv3'), sd, f'{p}.bn3')
    if f'{p}.downsample.0.weight' in sd:
        idt = _bn(_conv(x, sd, f'{p}.downsample.0', stride=stride), sd, f'{p}.downsample.1')
    return F.relu(out + idt)


def resnet_encoder(x, sd, p, kind):
    """models/resnet.py:265-290 with fused_initial=False (unet.py:584-587, fpn.py:233-236):
    body.0 = conv7x7 s2 + BN + ReLU, body.1 = Sequential(MaxPool(3, 2, 1), layer1), body.2..4 = layer2..4."""
    layers, bottle, groups, _ = RESNETS[kind]     # widths are read off the state_dict's weight shapes
    block = _bottleneck if bottle else _basic_block
    groups = groups if bottle else None
    feats = OrderedDict()
    x = F.relu(_bn(_conv(x, sd, f'{p}.0.0', stride=2, padding=3), sd, f'{p}.0.1'))
    feats['0'] = x
    for li, nblocks in enumerate(layers):
        if li == 0:
            x = F.max_pool2d(x, 3, 2, 1)
        for bi in range(nblocks):
            bp = f'{p}.1.1.{bi}' if li == 0 else f'{p}.{li + 1}.{bi}'
            stride = 2 if (li > 0 and bi == 0) else 1
            x = block(x, sd, bp, stride) if groups is None else block(x, sd, bp, stride, groups)
        feats[str(li + 1)] = x
    return feats


def unet_decoder(feats, sd, p, bridges, size):
    """models/unet.py:178-249.  ``bridges`` = log2 of the first encoder stride (0 for U22, 1 for ResNet encoders)."""
    names = list(feats.keys())
    x = list(feats.values())
    depth = len(x) - 1 + bridges
    last_inner = x[-1]
    results = [last_inner]
    for i in range(depth - 1, -1, -1):
        has_lat = (i - bridges) >= 0
        lateral = x[i - bridges] if has_lat else None
        if lateral is not None:
            top = F.interpolate(last_inner, size=lateral.shape[2:], mode='nearest')
        else:
            top = F.interpolate(last_inner, scale_factor=2, mode='nearest')
        if f'{p}.inner_blocks.{i}.weight' in sd:  # otherwise nn.Identity (unet.py:121-128)
            top = _conv(top, sd, f'{p}.inner_blocks.{i}')
        inp = torch.cat((lateral, top), 1) if lateral is not None else top  # cat_order 0 (unet.py:219-224)
        last_inner = unet_block(inp, sd, f'{p}.layer_blocks.{i}')
        results.insert(0, last_inner)
    final = F.interpolate(last_inner, size=size, mode='bilinear', align_corners=False)  # unet.py:237
    results.insert(0, final)
    names = ['out'] + names
    out = OrderedDict(zip(names, results))
    out.update(OrderedDict(('encoder.' + k, v) for k, v in feats.items()))
    return out


def fpn_decoder(feats, sd, p):
    """torchvision FeaturePyramidNetwork.forward with conv(+bias), no norm, no activation (fpn.py:79-134)."""
    names = list(feats.keys())
    x = list(feats.values())
    last_inner = _conv(x[-1], sd, f'{p}.inner_blocks.{len(x) - 1}.0')
    results = [_conv(last_inner, sd, f'{p}.layer_blocks.{len(x) - 1}.0', padding=1)]
    for idx in range(len(x) - 2, -1, -1):
        lat = _conv(x[idx], sd, f'{p}.inner_blocks.{idx}.0')
        top = F.interpolate(last_inner, size=lat.shape[-2:], mode='nearest')
        last_inner = lat + top
        results.insert(0, _conv(last_inner, sd, f'{p}.layer_blocks.{idx}.0', padding=1))
    names.append('pool')
    results.append(F.max_pool2d(results[-1], 1, 2, 0))  # fpn.py:67-76
    return OrderedDict(zip(names, results))


def read_out(x, sd, p, final=None, stride=1):
    """models/commons.py:461-511: conv kxk (bias, stride) -> BN -> ReLU -> Dropout2d (identity in eval) -> conv1x1 (bias).
    The kernel size (``kernel_size_<head>``, models/cpn.py:179-229) and the mid width (``*_head_channels``) are read off
    the state_dict."""
    k = sd[f'{p}.block.0.weight'].shape[-1]
    x = F.relu(_bn(_conv(x, sd, f'{p}.block.0', stride=stride, padding=k // 2), sd, f'{p}.block.1'))
    x = _conv(x, sd, f'{p}.block.4')
    if final is not None:
        x = final(x)
    return x


def backbone(x, sd, arch, p='core.backbone'):
    cfg = ARCHS[arch]
    xin = normalize(x)
    if cfg['encoder'] == 'unet':
        feats = unet_encoder(xin, sd, f'{p}.body')
        return unet_decoder(feats, sd, f'{p}.unet', bridges=0, size=x.shape[-2:])
    feats = resnet_encoder(xin, sd, f'{p}.body', cfg['encoder'])
    if cfg['decoder'] == 'unet':
        return unet_decoder(feats, sd, f'{p}.unet', bridges=1, size=x.shape[-2:])
    return fpn_decoder(feats, sd, f'{p}.fpn')


def _equal_size(x, ref_hw):
    """models/cpn.py:109-115"""
    if tuple(x.shape[2:]) != tuple(ref_hw):
        x = F.interpolate(x, tuple(ref_hw), mode='bilinear', align_corners=False)
    return x


CORE_KW = ('refinement_margin', 'contour_head_stride', 'refinement_head_stride', 'refinement_full_res')


def cpn_core5(x, sd, arch, refinement_margin=3., contour_head_stride=1, refinement_head_stride=1,
              refinement_full_res=True):
    """models/cpn.py:238-283 -> scores, locations, refinement, fourier, uncertainty (raw head tensors, NCHW fp32).
    The variants are read off the state_dict: score head width (classes > 2), refinement head width (2 * buckets),
    presence of ``core.uncertainty_head`` (4 sigmoid outputs, cpn.py:208-219)."""
    cfg = ARCHS[arch]
    feats = backbone(x, sd, arch)
    hf, rf = feats[cfg['head_key']], feats[cfg['ref_key']]
    cs = contour_head_stride
    scores = read_out(hf, sd, 'core.score_head', stride=cs)
    locations = read_out(hf, sd, 'core.location_head', stride=cs)
    fourier = read_out(hf, sd, 'core.fourier_head', stride=cs)
    uncertainty = None
    if 'core.uncertainty_head.block.0.weight' in sd:
        uncertainty = read_out(hf, sd, 'core.uncertainty_head', final=torch.sigmoid, stride=cs)
    if refinement_full_res:                       # models/cpn.py:277-278
        rf = _equal_size(rf, x.shape[2:])
    refinement = read_out(rf, sd, 'core.refinement_head', final=lambda t: torch.tanh(t) * refinement_margin + 0.,
                          stride=refinement_head_stride)
    refinement = _equal_size(refinement, x.shape[2:])
    return scores, locations, refinement, fourier, uncertainty


def cpn_core(x, sd, arch, refinement_margin=3., **core_kw):
    """``cpn_core5`` without the uncertainty map."""
    return cpn_core5(x, sd, arch, refinement_margin, **core_kw)[:4]


# ----------------------------------------------------------------------------------------------------------------------
# Post-head chain
# ----------------------------------------------------------------------------------------------------------------------

def fouriers2contours(fourier, locations, samples=64, sampling=None):
    """ops/cpn.py:44-95.  fourier [..., order, 4], locations [..., 2] -> ([..., samples, 2], sampling)."""
    fourier = torch.as_tensor(fourier)
    locations = torch.as_tensor(locations)
    order = fourier.shape[-2]
    sampling_ = sampling
    if sampling is None:
        sampling = sampling_ = torch.linspace(0, 1.0, samples)
    samples = sampling.shape[-1]
    sampling = sampling[..., None, :]
    c = float(np.pi) * 2 * (torch.arange(1, order + 1)[..., None]) * sampling
    c_cos, c_sin = torch.cos(c), torch.sin(c)
    con = torch.zeros(fourier.shape[:-2] + (samples, 2))
    con = con + locations[..., None, :]
    con += (fourier[..., None, (1, 3)] * c_sin[(...,) + (None,) * 1]).sum(-3)
    con += (fourier[..., None, (0, 2)] * c_cos[(...,) + (None,) * 1]).sum(-3)
    return con, sampling_


def box_area(b):
    return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])


def nms(boxes, scores, iou_threshold):
    """Restatement of ``torch.ops.torchvision.nms`` (torchvision 0.26 CPU kernel semantics):
    stable descending score order, suppress j when IoU(i, j) > thr (strict; NaN never suppresses), areas without +1,
    fp32 arithmetic.  Returns kept indices (int64) in descending score order."""
    boxes = np.asarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=np.float32).reshape(-1)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-scores, kind='stable')
    x1, y1, x2, y2 = (boxes[order, k] for k in range(4))
    areas = ((x2 - x1) * (y2 - y1)).astype(np.float32)
    thr = np.float32(iou_threshold)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    zero = np.float32(0)
    with np.errstate(invalid='ignore', divide='ignore'):
        for i in range(n):
            if suppressed[i]:
                continue
            keep.append(order[i])
            if i + 1 >= n:
                break
            xx1 = np.maximum(x1[i], x1[i + 1:])
            yy1 = np.maximum(y1[i], y1[i + 1:])
            xx2 = np.minimum(x2[i], x2[i + 1:])
            yy2 = np.minimum(y2[i], y2[i + 1:])
            w = np.maximum(zero, xx2 - xx1)
            h = np.maximum(zero, yy2 - yy1)
            inter = (w * h).astype(np.float32)
            ovr = inter / (areas[i] + areas[i + 1:] - inter)
            suppressed[i + 1:] |= ovr > thr
    return np.asarray(keep, dtype=np.int64)


def batched_box_nmsi(boxes, scores, iou_threshold, batch_size=None):
    """ops/cpn.py:189-227 (per-image NMS with the 50 000-chunk rule)."""
    batch_size = NMS_BATCH_SIZE if batch_size is None else batch_size
    keeps = []
    for con, sco in zip(boxes, scores):
        con, sco = np.asarray(con, np.float32), np.asarray(sco, np.float32)
        num = con.shape[0]
        if num <= batch_size:
            idx = nms(con, sco, iou_threshold)
        else:
            idx = np.zeros((0,), np.int64)
            for s in range(0, num, batch_size):
                e = min(s + batch_size, num)
                idx = np.concatenate((idx, nms(con[s:e], sco[s:e], iou_threshold) + s))
            if idx.size > 0:
                idx = idx[nms(con[idx], sco[idx], iou_threshold)]
        keeps.append(idx)
    return keeps


def refinement_bucket_weight(index, base_index):
    """ops/cpn.py:238-244"""
    dist = torch.abs(index + 0.5 - base_index)
    sel = dist > 1
    dist = 1. - dist
    dist[sel] = 0
    return dist


def resolve_refinement_buckets(samplings, num_buckets):
    """ops/cpn.py:247-255"""
    base_index = samplings * num_buckets
    base_index_int = base_index.long()
    a, b, c = base_index_int - 1, base_index_int, base_index_int + 1
    return ((a % num_buckets, refinement_bucket_weight(a, base_index)),
            (b % num_buckets, refinement_bucket_weight(b, base_index)),
            (c % num_buckets, refinement_bucket_weight(c, base_index)))


def cpn_post(scores, locations, refinement, fourier, original_size, order=5, samples=32, score_thresh=.9,
             nms_thresh=.2, refinement_iterations=4, offsets=None, nms_on=True, scores_lower_bound=None,
             scores_upper_bound=None, uncertainty=None, certainty_thresh=None, uncertainty_nms=False):
    """models/cpn.py:575-734 (eval mode) including the variants: score_channels > 2 (softmax / argmax, :583-585),
    uncertainty head (certainty filter :617-618, box_uncertainties :723-726, uncertainty_nms) and bucketed local
    refinement (refinement.shape[1] = 2 * buckets, :73-82).

    Inputs are the raw NCHW head tensors of ``cpn_core5``.  Returns an OrderedDict of per-image lists of tensors,
    keys as the reference: contours, boxes, scores, classes, locations, fourier, contour_proposals, box_uncertainties.
    """
    n = scores.shape[0]
    H, W = original_size
    score_channels = scores.shape[1]

    def bounds(sc):  # cpn.py:118-123
        if scores_upper_bound is not None:
            sc = torch.minimum(sc, _equal_size(scores_upper_bound, sc.shape[2:]))
        if scores_lower_bound is not None:
            sc = torch.maximum(sc, _equal_size(scores_lower_bound, sc.shape[2:]))
        return sc
    if score_channels == 1:
        sc = bounds(torch.sigmoid(scores))
        classes = torch.squeeze((sc > score_thresh).long(), 1)
    elif score_channels == 2:
        sc = bounds(F.softmax(scores, dim=1)[:, 1:2])
        classes = torch.squeeze((sc > score_thresh).long(), 1)
    else:
        sc = bounds(F.softmax(scores, dim=1))
        classes = torch.argmax(sc, dim=1).long()
    h, w = fourier.shape[-2:]
    fo = fourier.view(n, fourier.shape[1] // 4, 4, h, w)
    if order < fo.shape[1]:
        fo = fo[:, :order]
    # rel -> abs locations (ops/cpn.py:15-41): channel 0 += x index, channel 1 += y index
    grid = torch.stack((torch.arange(w)[None] + torch.zeros(h)[:, None],
                        torch.zeros(w)[None] + torch.arange(h)[:, None]), 0)
    loc = locations + grid
    fg = classes > 0
    if certainty_thresh is not None and uncertainty is not None:
        fg = fg & (uncertainty.mean(1) < (1 - certainty_thresh))
    b, y, x = torch.where(fg)
    sel_fourier = fo[b, :, :, y, x]
    sel_loc = loc[b, :, y, x]
    sel_classes = classes[b, y, x]
    sel_scores = sc[b, 0, y, x] if score_channels in (1, 2) else sc[b, sel_classes, y, x]
    sel_unc = uncertainty[b, :, y, x] if uncertainty is not None else None
    proposals, sampling = fouriers2contours(sel_fourier, sel_loc, samples=samples)
    scale = (torch.as_tensor((H, W), dtype=torch.float) / torch.as_tensor((h, w), dtype=torch.float)).flip(-1)
    proposals = proposals * scale  # ops/cpn.py:98-127
    sel_fourier = sel_fourier.clone()
    sel_fourier[..., [0, 1]] = sel_fourier[..., [0, 1]] * scale[0]  # ops/cpn.py:133-137
    sel_fourier[..., [2, 3]] = sel_fourier[..., [2, 3]] * scale[1]
    sel_loc = sel_loc * scale
    if refinement is not None and refinement_iterations > 0:  # cpn.py:63-85
        num_buckets = refinement.shape[1] // 2
        det = proposals
        for _ in range(refinement_iterations):
            det = torch.round(det)
            det[..., 0].clamp_(0, W - 1)
            det[..., 1].clamp_(0, H - 1)
            idx = det.long()
            if num_buckets == 1:
                responses = refinement[b[:, None], :, idx[:, :, 1], idx[:, :, 0]]
            else:
                responses = None
                for bucket_indices, bucket_weights in resolve_refinement_buckets(sampling, num_buckets):
                    bckt_idx = torch.stack((bucket_indices * 2, bucket_indices * 2 + 1), -1)
                    cur = refinement[b[:, None, None], bckt_idx, idx[:, :, 1, None], idx[:, :, 0, None]]
                    cur = cur * bucket_weights[..., None]
                    responses = cur if responses is None else responses + cur
            det = det + responses
        contours = det
    else:
        contours = proposals
    contours[..., 0].clamp_(0, W - 1)  # cpn.py:661-663 (in place: aliases proposals when refinement is off)
    contours[..., 1].clamp_(0, H - 1)
    if contours.numel() > 0:
        boxes = torch.cat((contours.min(1).values, contours.max(1).values), 1)
    else:
        boxes = torch.empty((0, 4))
    if offsets is not None:  # cpn.py:695-702
        off = torch.as_tensor(offsets, dtype=torch.float)[b]
        contours += off[:, None]
        proposals += off[:, None]
        boxes += off.repeat((1, 2))
        sel_loc += off
    outputs = OrderedDict(contours=contours, boxes=boxes, scores=sel_scores, classes=sel_classes, locations=sel_loc,
                          fourier=sel_fourier, contour_proposals=proposals)
    if sel_unc is not None:
        outputs['box_uncertainties'] = sel_unc
    per_image = OrderedDict((k, [v[b == i] for i in range(n)]) for k, v in outputs.items())  # cpn.py:42-50
    if nms_on:
        if uncertainty_nms and sel_unc is not None:  # cpn.py:723-726
            nms_w = [s * (1. - u.mean(1)) for s, u in zip(per_image['scores'], per_image['box_uncertainties'])]
        else:
            nms_w = per_image['scores']
        keeps = batched_box_nmsi([t.numpy() for t in per_image['boxes']], [t.numpy() for t in nms_w], nms_thresh)
        per_image = OrderedDict((k, [v[i][torch.as_tensor(keeps[i])] for i in range(n)])
                                for k, v in per_image.items())
    per_image.setdefault('box_uncertainties', None)
    return per_image


def cpn_forward(x, sd, arch, **kw):
    """``model(x)`` of the reference in eval mode: core + post chain.  ``kw`` are the mutable CPN attributes."""
    x = torch.as_tensor(x, dtype=torch.float32)
    core_kw = {k: kw.pop(k) for k in CORE_KW if k in kw}      # constructor options that shape the core
    with torch.no_grad():
        scores, locations, refinement, fourier, uncertainty = cpn_core5(x, sd, arch, **core_kw)
        return cpn_post(scores, locations, refinement, fourier, x.shape[-2:], uncertainty=uncertainty, **kw)


# ----------------------------------------------------------------------------------------------------------------------
# Tiling driver (celldetection_scripts/cpn_inference.py:311-429, util/util.py:1305-1354)
# ----------------------------------------------------------------------------------------------------------------------

def get_tiling_slices(size, crop_size, strides):
    """util/util.py:1305-1354.  Returns (slices per tile as tuples of python slices, overlaps [n, dims, 2], grid)."""
    import itertools
    assert len(size) == len(crop_size) == len(strides)
    slices, overlaps, shape = [], [], []
    for axis in range(len(size)):
        sz, cr, st = int(size[axis]), int(crop_size[axis]), int(strides[axis])
        if cr >= sz:
            tl = [sz]
        else:
            tl = list(range(cr, 1 + cr + int(math.ceil((sz - cr) / st)) * st, st))
        stops = [min(t, sz) for t in tl]
        starts = [max(0, s - cr) for s in stops]
        ov_start = [(starts[0] if j == 0 else stops[j - 1]) - starts[j] for j in range(len(starts))]
        ov_end = ov_start[1:] + [0]
        slices.append([slice(a, b_) for a, b_ in zip(starts, stops)])
        overlaps.append(list(zip(ov_start, ov_end)))
        shape.append(len(starts))
    tiles = list(itertools.product(*slices))
    ovs = np.array(list(itertools.product(*overlaps)), dtype=np.int64).reshape(len(tiles), len(size), 2)
    return tiles, ovs, tuple(shape)


def remove_border_contours(contours, size, padding=1, top=True, right=True, bottom=True, left=True, offsets=None):
    """ops/cpn.py:258-290 -> boolean keep mask."""
    h, w = size[:2]
    contours = torch.as_tensor(contours)
    if offsets is not None:
        contours = contours + offsets
    x, y = contours[..., 0], contours[..., 1]
    keep = torch.ones(len(contours), dtype=torch.bool)
    if top:
        keep = keep & (y > padding).all(1)
    if right:
        keep = keep & (x < (w - padding)).all(1)
    if bottom:
        keep = keep & (y < (h - padding)).all(1)
    if left:
        keep = keep & (x > padding).all(1)
    return keep


def filter_contours_by_stitching_rule(contours, tile_size, overlaps, rule='ex_br', offsets=None):
    """ops/cpn.py:293-325 ('ex_br'): drop contours whose every vertex lies in the right/bottom overlap of the tile."""
    contours = torch.as_tensor(contours)
    tile_size = torch.as_tensor(tile_size)
    overlaps = torch.as_tensor(overlaps)
    if offsets is not None:
        contours = contours + offsets
    assert 'ex_br' in rule.split(',')
    stop = (tile_size - overlaps[:, 1])[[1, 0]]
    return ~((contours >= stop).any(-1).all(-1))


def tta_example_transform(crop, rep_idx):
    """A test-time transform in TileLoader's protocol (cpn_inference.py:118): repetition 0 sees the crop itself, repetition 1 its
    horizontal mirror image, repetition 2 the vertical one with rotated channels (the reference never maps detections back, so
    the mirrored detections are new ones).  Shared by oracle/make_golden.py and the tests."""
    if rep_idx == 0:
        return crop, None
    out = crop[:, ::-1] if rep_idx % 2 else np.roll(crop[::-1], 1, axis=-1)
    return np.ascontiguousarray(out), dict(rep=rep_idx)


def apply_model(img, sd, arch, crop_size, strides, border_removal=4, batch_size=1, mask=None, point_mask=None,
                point_mask_exclusive=False, reps=1, transforms=None, **kw):
    """cpn_inference.py:311-411 for one model, ``stitching_rule='nms'``.

    ``img`` is uint8 or float HxWx3.  uint8 tiles become float/255 (lightning_base.py:774-780).  Per tile: model with
    ``offsets=[w0, h0]`` (and, with ``mask`` / ``point_mask``, the crop as upper / lower score bound; tiles whose crop
    is empty are skipped, TileLoader :93-111); border removal in tile-local coordinates with sides disabled at the
    image border (cpn_inference.py:370-387); then concat and one global NMS with the model's ``nms_thresh`` (:405-408).
    ``reps`` / ``transforms``: every tile is inferred ``reps`` times, ``transforms(crop, rep_idx)`` changing the input crop only
    (TileLoader :85-91,114-118; the meta it returns is never used to map detections back).
    """
    img = np.asarray(img)
    H, W = img.shape[:2]
    crop = (min(crop_size, H), min(crop_size, W)) if np.isscalar(crop_size) else tuple(crop_size)
    strd = (strides, strides) if np.isscalar(strides) else tuple(strides)
    tiles, _, grid = get_tiling_slices((H, W), crop, strd)
    nms_thresh = kw.get('nms_thresh', .2)
    acc = OrderedDict()
    for t_idx, (sl_h, sl_w) in enumerate(tiles):
        gy, gx = np.unravel_index(t_idx, grid)
        bkw = {}
        if mask is not None:
            mc = np.asarray(mask)[sl_h, sl_w]
            if not np.any(mc):
                continue
            bkw['scores_upper_bound'] = torch.as_tensor(mc.astype('float32'))[None, None]
        if point_mask is not None:
            pc = np.asarray(point_mask)[sl_h, sl_w]
            if not np.any(pc):
                continue
            bkw['scores_lower_bound'] = torch.as_tensor(np.clip(pc, 0., 1.).astype('float32'))[None, None]
            if point_mask_exclusive:
                bkw['scores_upper_bound'] = bkw['scores_lower_bound']
        for rep_idx in range(reps):
            crop_img = img[sl_h, sl_w]
            if transforms is not None:
                assert mask is None and point_mask is None       # cpn_inference.py:116-117
                crop_img, _ = transforms(crop_img, rep_idx)
            x = torch.as_tensor(np.ascontiguousarray(crop_img)).permute(2, 0, 1)[None]
            x = x.float() / 255 if crop_img.dtype == np.uint8 else x.float()
            off = torch.as_tensor([[sl_w.start, sl_h.start]], dtype=torch.float)
            out = cpn_forward(x, sd, arch, offsets=off, **bkw, **kw)
            con = out['contours'][0]
            keep = remove_border_contours(con, crop_img.shape[:2], border_removal, top=gy > 0, right=gx < grid[1] - 1,
                                          bottom=gy < grid[0] - 1, left=gx > 0, offsets=-off[0])
            for k, v in out.items():
                if v is None:
                    continue
                acc.setdefault(k, []).append(v[0][keep])
    res = OrderedDict((k, torch.cat(v, 0)) for k, v in acc.items())
    keep = torch.as_tensor(nms(res['boxes'].numpy(), res['scores'].numpy(), nms_thresh))
    return OrderedDict((k, v[keep]) for k, v in res.items())


def box_iou(a, b):
    """torchvision.ops.boxes.box_iou (third party): inter / (area_a + area_b - inter), fp32."""
    a, b = torch.as_tensor(a, dtype=torch.float32), torch.as_tensor(b, dtype=torch.float32)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None] - inter)


def filter_by_box_voting(boxes, thresh, min_vote):
    """ops/boxes.py:53-83: votes = (iou * (iou > thresh)).sum(-1); keep votes >= min_vote.  Returns (keep, votes[keep])."""
    iou = box_iou(boxes, boxes)
    iou = iou * (iou > thresh)
    votes = iou.sum(-1)
    m = votes >= min_vote
    return torch.arange(len(votes))[m], votes[m]


def apply_models(img, sds, archs, crop_size, strides, min_vote=1, **kw):
    """cpn_inference.py:354-429 for a list of models: per-model tiled inference + stitch NMS, concatenation, box voting
    (``min_vote > 1``) and the final NMS with the (last) model's ``nms_thresh`` (:417-427)."""
    assert len(sds) >= min_vote >= 1
    results = None
    for sd, arch in zip(sds, archs):
        r = apply_model(img, sd, arch, crop_size, strides, **kw)
        results = r if results is None else OrderedDict((k, torch.cat((results[k], r[k]), 0)) for k in results)
    nms_thresh = kw.get('nms_thresh', .2)
    if len(sds) > 1 and len(results['scores']):
        if min_vote > 1:
            keep, votes = filter_by_box_voting(results['boxes'], nms_thresh, min_vote)
            results = OrderedDict((k, v[keep]) for k, v in results.items())
            results['votes'] = votes
        keep = torch.as_tensor(nms(results['boxes'].numpy(), results['scores'].numpy(), nms_thresh))
        results = OrderedDict((k, v[keep]) for k, v in results.items())
    return results
