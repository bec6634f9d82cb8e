"""Shared helpers of the test-suite (fixtures -> state dicts, matching utilities)."""
import json
import os
from collections import OrderedDict

import numpy as np
import torch

from conftest import GOLDEN
from celldetection_b200.utils.synth import synth_state_dict

CALIB_KEYS = [f'core.{h}_head.block.4.{p}' for h in ('score', 'fourier', 'location', 'refinement')
              for p in ('weight', 'bias')]
MODEL_FIXTURES = ['model_cpnu22_n1_128', 'model_cpnresnet18fpn_n2_128', 'model_cpnresnext101unet_n1_128',
                  'model_cpnu22_n2_96x160_s64']


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def key_spec(arch):
    with open(os.path.join(GOLDEN, 'state_dict_keys.json')) as f:
        keys = json.load(f)[arch]
    return OrderedDict((k, tuple(s)) for k, s in keys)


VARIANT_FIXTURES = ['model_cpnu22_c4_unc_b6', 'model_cpnresnext101unet_unc_b4', 'model_cpnresnet18fpn_c3_b2',
                    # shape-changing head options: per-head kernel sizes / mid widths / strides / fpn_channels
                    'model_cpnu22_k357_mid64', 'model_cpnresnet18fpn_mid_stride2']


def fixture_ctor(z):
    """(constructor kwargs, attributes set after construction) of a variant fixture ({} / {} for the default models)."""
    if 'ctor' not in z.files:
        return {}, {}
    return json.loads(str(z['ctor'])), json.loads(str(z['attrs']))


def fixture_state_dict(z, arch, seed):
    """Rebuild the exact state_dict a fixture was minted with: seeded synthetic weights + stored calibrated heads."""
    sd = synth_state_dict(key_spec(str(z['spec_key']) if 'spec_key' in z.files else arch), seed=seed)
    for f in z.files:
        if f.startswith('calib/'):
            sd[f[len('calib/'):]] = torch.from_numpy(np.array(z[f]))
    return sd


def ensemble_state_dicts(z):
    """The two CpnU22 members of the ensemble fixture (same seed, different calibrated / perturbed heads)."""
    sds = []
    for j, seed in enumerate(int(v) for v in z['seeds']):
        sd = synth_state_dict(key_spec('CpnU22'), seed=seed)
        for f in z.files:
            if f.startswith(f'calib{j}/'):
                sd[f[len(f'calib{j}/'):]] = torch.from_numpy(np.array(z[f]))
        sds.append(sd)
    return sds


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return 0.
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def match_by_box(boxes_a, boxes_b, max_dist=2.0):
    """Greedy one-to-one matching of detections by box distance (max |coordinate difference| <= max_dist px); robust
    to near-tie score ordering differences and to degenerate (zero-area) boxes.  Returns index pairs (ia, ib)."""
    boxes_a, boxes_b = np.asarray(boxes_a, np.float64).reshape(-1, 4), np.asarray(boxes_b, np.float64).reshape(-1, 4)
    pairs, used = [], set()
    for i, a in enumerate(boxes_a):
        best, bj = max_dist, -1
        for j, b in enumerate(boxes_b):
            if j in used:
                continue
            d = np.abs(a - b).max()
            if d <= best:
                best, bj = d, j
        if bj >= 0:
            used.add(bj)
            pairs.append((i, bj))
    return pairs
