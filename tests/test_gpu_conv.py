"""GPU: per-layer-type convolution parity through the C ABI's cpn_conv2d, against
torch.nn.functional.conv2d on the CPU in fp32.  Each engine runs in a child process with a timeout."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
TC_CASES = list(range(0, 16))
ALL_CASES = list(range(0, 18))


def _run(engine, ids, timeout=600):
    p = subprocess.run([sys.executable, os.path.join(HERE, 'gpu_conv_check.py'), engine] + [str(i) for i in ids],
                       capture_output=True, text=True, timeout=timeout)
    rows = [json.loads(l) for l in p.stdout.splitlines() if l.startswith('{')]
    assert p.returncode == 0 and len(rows) == len(ids), (p.returncode, p.stdout[-2000:], p.stderr[-2000:])
    return rows


@pytest.mark.gpu
def test_conv_simt_fp32_matches_torch():
    for r in _run('simt32', ALL_CASES):
        assert 'error' not in r, r
        assert r['rel_err'] < 2e-5 and not r['nan'], r          # fp32, different summation order only


@pytest.mark.gpu
def test_conv_simt_fp16_storage_matches_torch():
    for r in _run('simt16', ALL_CASES):
        assert 'error' not in r, r
        assert r['rel_err'] < 2e-3 and not r['nan'], r          # fp16 output rounding (2^-11 relative)


@pytest.mark.gpu
def test_conv_tcgen05_matches_torch():
    for r in _run('tcgen05', TC_CASES):
        assert 'error' not in r, r
        assert r['rel_err'] < 2e-3 and not r['nan'], r          # fp16 operands, fp32 accumulate, fp16 store


@pytest.mark.gpu
def test_conv_tcgen05_split_fp16x3_matches_fp32():
    """Three-pass split-fp16 engine on unrounded fp32 operands.  Operand rounding is gone (~1e-6 for K <= 1e3); what
    remains grows linearly with K (~5e-9 * K, 1.4e-4 at K = 27 648): the tensor core's fp32 accumulator truncates."""
    for r in _run('tcgen05x3', TC_CASES):
        assert 'error' not in r, r
        cin, cout, k, stride, groups = r['shape'][:5]
        kk = (cin // groups) * k * k if groups == 1 else 64 * k * k
        assert r['rel_err'] < max(5e-6, 1e-8 * kk) and not r['nan'], r


@pytest.mark.gpu
def test_conv_tcgen05_fp16_plus_e4m3_corrections_matches_fp32():
    """Two-pass engine (one kind::f16 pass + one kind::f8f6f4 correction pass over the e4m3 residual operands, same fp32
    accumulator) on unrounded fp32 operands.  The corrections carry 4 significant bits, so the operand error falls from
    2^-11 (single-pass fp16, ~3e-4 per layer) to ~2^-15 per element; what remains averages over K."""
    for r in _run('tcgen05f8', TC_CASES):
        assert 'error' not in r, r
        assert r['rel_err'] < 6e-5 and not r['nan'], r
