"""Deterministic synthetic weights for benchmarks and parity tests (there is no network for checkpoints).

``synth_state_dict`` draws a well-conditioned "trained-like" state_dict for any key/shape specification (He-normal
convolutions, perturbed BatchNorm statistics, damped residual branches) so that activations keep O(1) magnitude through
100+ layers -- the reference's own default init collapses to ~1e-4 logits and zero detections.
``calibrate_heads_`` is the head-calibration recipe of SURVEY.md 8(d): rescale the final 1x1 convolutions of the
score / fourier / location heads from one measured forward so that a stated fraction of pixels becomes proposals.
Both reference and B200 model then load the *same* state_dict.
"""
import math
from collections import OrderedDict

import torch


def synth_state_dict(spec, seed=0, dtype=torch.float32):
    """spec: OrderedDict key -> shape (torch.Size/tuple) or (shape, role).  Roles are inferred from key names."""
    g = torch.Generator().manual_seed(int(seed))
    sd = OrderedDict()
    for key, val in spec.items():
        shape = tuple(val[0]) if (isinstance(val, tuple) and len(val) == 2 and isinstance(val[1], str)) else tuple(val)
        leaf = key.rsplit('.', 1)[-1]
        if key == 'order_weights':
            order = shape[0]
            x = torch.arange(order).float()
            sd[key] = (1 + 4 * (1 - (x / max(order - 1, 1)).clamp(0., 1.)) ** 2)[:, None]
        elif leaf == 'num_batches_tracked':
            sd[key] = torch.tensor(0, dtype=torch.long)
        elif leaf == 'running_mean':
            sd[key] = torch.randn(shape, generator=g) * 0.05
        elif leaf == 'running_var':
            sd[key] = torch.rand(shape, generator=g) * 0.4 + 0.8
        elif len(shape) == 4:                                  # conv weight: He normal
            fan_in = shape[1] * shape[2] * shape[3]
            sd[key] = torch.randn(shape, generator=g) * math.sqrt(2. / fan_in)
        elif leaf == 'weight':                                 # BatchNorm gamma
            gamma = torch.rand(shape, generator=g) * 0.4 + 0.8
            if key.endswith('.bn3.weight') or (key.endswith('.bn2.weight') and _is_basic(spec, key)):
                gamma = gamma * 0.25                           # last BN of a residual branch
            sd[key] = gamma
        elif leaf == 'bias':
            sd[key] = torch.randn(shape, generator=g) * 0.05
        else:
            raise KeyError(f'cannot synthesise {key} {shape}')
        if sd[key].is_floating_point():
            sd[key] = sd[key].to(dtype)
    return sd


def _is_basic(spec, key):
    """bn2 is the last BN of a BasicBlock iff the block has no conv3."""
    prefix = key[:-len('bn2.weight')]
    return (prefix + 'conv3.weight') not in spec


def _mean_std(t):
    """Mean / unbiased std in float64 with numpy's single-threaded pairwise summation: torch's CPU reductions split the
    work by the number of OpenMP threads, so their last bits -- and with them the calibrated weights and every digest
    derived from them -- would depend on the host's core count (seen as a 1-GPU-box vs 2-GPU-box digest mismatch)."""
    a = t.detach().double().cpu().numpy().ravel()
    return float(a.mean()), float(a.std(ddof=1)) if a.size > 1 else 0.


@torch.no_grad()
def calibrate_heads_(sd, core_fn, x, fg_fraction=0.02, score_std=2., fourier_std=3., location_std=1.,
                     score_thresh=0.9, refinement_std=0.5, refinement_margin=3., uncertainty_std=1.5):
    """In place: rescale ``core.{score,fourier,location}_head.block.4`` so that on calibration input ``x`` the raw score
    logits have std ``score_std`` and mean such that ``fg_fraction`` of the pixels exceed ``score_thresh`` (normal
    approximation), fourier std -> ``fourier_std`` px, location std -> ``location_std`` px.

    If ``core_fn`` also returns ``refinement`` (the 3*tanh output), the refinement head's final 1x1 convolution is
    rescaled (two passes, through atanh) so that its pre-activation has std ``refinement_std``: an uncalibrated head
    saturates tanh almost everywhere, which is unlike trained weights and makes the output needlessly steep.

    ``core_fn(x, sd) -> dict(scores [N,C,h,w], locations [N,2,h,w], fourier [N,4*order,h,w][, refinement]
    [, uncertainty])`` is one forward of whichever implementation is at hand (an ``uncertainty`` entry, the sigmoid output
    of the uncertainty head, gets its logits calibrated to std ``uncertainty_std``).  Returns the measured pre-calibration statistics.
    """
    out = core_fn(x, sd)
    stats = {}
    from statistics import NormalDist
    z = NormalDist().inv_cdf(1. - fg_fraction)
    target_mu = math.log(score_thresh / (1. - score_thresh)) - score_std * z
    for name, key, tstd, tmu in (('scores', 'core.score_head.block.4', score_std, target_mu),
                                 ('fourier', 'core.fourier_head.block.4', fourier_std, 0.),
                                 ('locations', 'core.location_head.block.4', location_std, 0.)):
        mu, std = _mean_std(out[name])
        stats[name] = (mu, std)
        s = tstd / max(std, 1e-12)
        dev, dt = sd[key + '.weight'].device, sd[key + '.weight'].dtype
        sd[key + '.weight'] = (sd[key + '.weight'].float() * s).to(dev, dt)
        sd[key + '.bias'] = ((sd[key + '.bias'].float() - mu) * s + tmu).to(dev, dt)
    key = 'core.uncertainty_head.block.4'
    if out.get('uncertainty') is not None and uncertainty_std:      # sigmoid outputs: calibrate the logits to N(0, std)
        u = out['uncertainty'].float().clamp(1e-6, 1 - 1e-6)
        zed = torch.log(u / (1 - u))
        mu, std = _mean_std(zed)
        stats['uncertainty_pre'] = (mu, std)
        s = uncertainty_std / max(std, 1e-12)
        dev, dt = sd[key + '.weight'].device, sd[key + '.weight'].dtype
        sd[key + '.weight'] = (sd[key + '.weight'].float() * s).to(dev, dt)
        sd[key + '.bias'] = ((sd[key + '.bias'].float() - mu) * s).to(dev, dt)
    key = 'core.refinement_head.block.4'
    if out.get('refinement') is not None and refinement_std:
        for it in range(2):
            r = (out if it == 0 else core_fn(x, sd))['refinement'].float()
            zed = torch.atanh((r / refinement_margin).clamp(-0.999, 0.999))
            mu, std = _mean_std(zed)
            stats[f'refinement_pre{it}'] = (mu, std)
            s = refinement_std / max(std, 1e-12)
            dev, dt = sd[key + '.weight'].device, sd[key + '.weight'].dtype
            sd[key + '.weight'] = (sd[key + '.weight'].float() * s).to(dev, dt)
            sd[key + '.bias'] = ((sd[key + '.bias'].float() - mu) * s).to(dev, dt)
    return stats
