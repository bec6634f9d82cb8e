"""Shared helpers of the test-suite (fixtures -> state dicts, matching utilities)."""
import json
import os
from collections import OrderedDict

import numpy as np
import torch

from conftest import GOLDEN
from celldetection_b200.utils.synth import synth_state_dict

CALIB_KEYS = [f'core.{h}_head.block.4.{p}' for h in ('score', 'fourier', 'location') for p in ('weight', 'bias')]
MODEL_FIXTURES = ['model_cpnu22_n1_128', 'model_cpnresnet18fpn_n2_128', 'model_cpnresnext101unet_n1_128',
                  'model_cpnu22_n2_96x160_s64']


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def key_spec(arch):
    with open(os.path.join(GOLDEN, 'state_dict_keys.json')) as f:
        keys = json.load(f)[arch]
    return OrderedDict((k, tuple(s)) for k, s in keys)


def fixture_state_dict(z, arch, seed):
    """Rebuild the exact state_dict a fixture was minted with: seeded synthetic weights + stored calibrated heads."""
    sd = synth_state_dict(key_spec(arch), seed=seed)
    for k in CALIB_KEYS:
        sd[k] = torch.from_numpy(np.array(z['calib/' + k]))
    return sd


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return 0.
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def match_by_box(boxes_a, boxes_b):
    """Greedy one-to-one matching of detections by box IoU (near-tie score ordering may differ between
    implementations).  Returns index pairs (ia, ib)."""
    boxes_a, boxes_b = np.asarray(boxes_a, np.float64), np.asarray(boxes_b, np.float64)
    pairs, used = [], set()
    for i, a in enumerate(boxes_a):
        best, bj = 0., -1
        for j, b in enumerate(boxes_b):
            if j in used:
                continue
            iw = max(0., min(a[2], b[2]) - max(a[0], b[0]))
            ih = max(0., min(a[3], b[3]) - max(a[1], b[1]))
            inter = iw * ih
            ua = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
            iou = inter / ua if ua > 0 else 0.
            if iou > best:
                best, bj = iou, j
        if bj >= 0 and best > 0.5:
            used.add(bj)
            pairs.append((i, bj))
    return pairs
