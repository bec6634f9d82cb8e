"""CPU: why the 1e-3 tensor gate of BASELINE.json's north_star needs more than fp16 operands everywhere.

The oracle is run in fp32 except for ONE layer -- the 7x7 convolution of every ReadOut head -- whose activations and
BN-folded weights are rounded to fp16 (products and sums stay exact, like a tensor core with fp32 accumulation).  That
single layer already moves the head tensors by more than 1e-3 (relative, ||a-b||inf / ||b||inf) on the flagship
architecture, so no mix of single-pass fp16 layers can meet the gate; the product therefore offers the fp16x3 engine
(hi/lo operand pairs, three tensor-core passes) as the parity engine and reports the single-pass fp16 engine's measured
deviation instead of claiming parity for it."""
import torch
import torch.nn.functional as F

import cpn_oracle as orc
from helpers import load_npz, fixture_state_dict


def _core_with_fp16_heads(x, sd, arch, round_act=True, round_w=True):
    orig = orc.read_out

    def ro(xx, sd_, p, final=None, stride=1):
        assert stride == 1
        k = sd_[f'{p}.block.0.weight'].shape[-1]
        g = sd_[f'{p}.block.1.weight'] / torch.sqrt(sd_[f'{p}.block.1.running_var'] + 1e-5)
        w = sd_[f'{p}.block.0.weight'] * g[:, None, None, None]
        b = (sd_[f'{p}.block.0.bias'] - sd_[f'{p}.block.1.running_mean']) * g + sd_[f'{p}.block.1.bias']
        if round_w:
            w = w.half().float()
        if round_act:
            xx = xx.half().float()
        y = F.relu(F.conv2d(xx.double(), w.double(), b.double(), padding=k // 2)).float()
        y = orc._conv(y, sd_, f'{p}.block.4')
        return final(y) if final is not None else y
    orc.read_out = ro
    try:
        return orc.cpn_core(x, sd, arch)
    finally:
        orc.read_out = orig


def test_one_fp16_layer_exceeds_the_1e3_gate():
    z = load_npz('model_cpnresnext101unet_n1_128')
    arch = str(z['arch'])
    seed = int(z['meta'][3])
    sd = fixture_state_dict(z, arch, seed)
    x = torch.from_numpy(z['x'])
    torch.set_num_threads(8)
    with torch.no_grad():
        ref = orc.cpn_core(x, sd, arch)
        both = _core_with_fp16_heads(x, sd, arch)
        w_only = _core_with_fp16_heads(x, sd, arch, round_act=False)
    err = lambda a, b: float((a - b).abs().max() / b.abs().max())  # noqa: E731
    errs = [err(a, b) for a, b in zip(both, ref)]          # scores, locations, refinement, fourier
    errs_w = [err(a, b) for a, b in zip(w_only, ref)]
    assert max(errs) > 1e-3, errs                          # measured: 1.2e-3, 1.1e-3, 1.7e-3, 4.4e-4
    assert max(errs_w) > 5e-4, errs_w                      # weight rounding alone: 6e-4 ... 1.1e-3
    assert min(errs) > 1e-4


def test_e4m3_correction_passes_would_keep_the_gate():
    """Design input for the next parity engine: emulating fp16 main pass + e4m3 correction
    passes with data-independent activation scales in EVERY layer of the oracle keeps the head tensors of CpnU22 within
    3e-4 of fp32, while single-pass fp16 is beyond 1e-3 (tests/study_fp8_corrections.py has the flagship numbers)."""
    import study_fp8_corrections as st
    z = load_npz('model_cpnu22_n1_128')
    arch = str(z['arch'])
    sd = fixture_state_dict(z, arch, int(z['meta'][3]))
    x = torch.from_numpy(z['x'])
    torch.set_num_threads(8)
    orig = orc._conv
    outs = {}
    try:
        for mode in ('fp32', 'fp16', 'e4m3fixed'):
            orc._conv = st.make_conv(mode)
            with torch.no_grad():
                outs[mode] = orc.cpn_core(x, sd, arch)
    finally:
        orc._conv = orig
    err = lambda m: max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs[m], outs['fp32']))  # noqa: E731
    assert err('fp16') > 1e-3
    assert err('e4m3fixed') < 3e-4
