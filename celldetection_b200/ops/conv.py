"""Stand-alone access to the plan's convolution kernels (per-layer parity tests and micro-benchmarks).

``conv2d`` takes/returns NCHW tensors like ``F.conv2d`` but runs ``cpn_conv2d``: NHWC staging, BatchNorm-free folded
weights, optional residual and ReLU in the epilogue.  ``engine='tcgen05'`` needs fp16-representable operands
(C_in per slab and C_out multiples of 64); ``engine='simt'`` is the strict fp32 CUDA-core kernel.
"""
import torch

from .. import _lib as L
from ..models.plan import expand_grouped, slab_of

__all__ = ['conv2d']


def conv2d(x, weight, bias=None, stride=1, padding=0, groups=1, residual=None, relu=False, engine='simt',
           half=None):
    """x [N,Cin,H,W], weight [Cout,Cin/groups,k,k] (CUDA) -> [N,Cout,Ho,Wo] fp32."""
    if not x.is_cuda:
        raise RuntimeError('conv2d runs on CUDA tensors only')
    lib = L.load()
    tc = engine == 'tcgen05'
    half = tc if half is None else half
    act_dt, tdt, es = (L.DT_F16, torch.float16, 2) if half else (L.DT_F32, torch.float32, 4)
    n, cin, h, w = x.shape
    cout, _, k, _ = weight.shape
    ho, wo = (h + 2 * padding - k) // stride + 1, (w + 2 * padding - k) // stride + 1
    pitch_in = cin if cin % 4 == 0 else (cin + 3) // 4 * 4
    xs = torch.zeros((n, h, w, pitch_in), dtype=tdt, device=x.device)
    xs[..., :cin] = x.permute(0, 2, 3, 1).to(tdt)
    out = torch.zeros((n, ho, wo, cout), dtype=tdt, device=x.device)
    wf = weight.float()
    if groups > 1:
        wf = expand_grouped(wf, groups)
    kslab, mode = slab_of(cin, cout, groups)
    if tc:
        wp = wf.permute(2, 3, 0, 1).reshape(k * k, cout, kslab).contiguous().half()
    else:
        wp = wf.permute(2, 3, 1, 0).reshape(k * k, kslab, cout).contiguous().float()
    wb = wp.view(torch.uint8).reshape(-1)
    b_off = (wb.numel() + 255) // 256 * 256
    bb = (bias if bias is not None else torch.zeros(cout, device=x.device)).float().contiguous().view(torch.uint8)
    blob = torch.zeros(b_off + bb.numel() + 256, dtype=torch.uint8, device=x.device)
    blob[:wb.numel()] = wb
    blob[b_off:b_off + bb.numel()] = bb.reshape(-1)
    op = L.Op()
    op.kind, op.engine = L.OP_CONV, (L.ENGINE_TCGEN05 if tc else L.ENGINE_SIMT)
    for v, (c, hh, ww, p) in ((op.src, (cin, h, w, pitch_in)), (op.dst, (cout, ho, wo, cout))):
        v.offset, v.n, v.h, v.w, v.c, v.pitch, v.dtype = 0, n, hh, ww, c, p, act_dt
    rs = None
    if residual is not None:
        rs = residual.permute(0, 2, 3, 1).to(tdt).contiguous()
        v = op.res
        v.offset, v.n, v.h, v.w, v.c, v.pitch, v.dtype = 0, n, residual.shape[2], residual.shape[3], cout, cout, act_dt
    op.w_offset, op.b_offset = 0, b_off
    op.r = op.s = k
    op.stride, op.pad, op.kslab, op.slab_mode = stride, padding, kslab, mode
    op.act = L.ACT_RELU if relu else L.ACT_NONE
    op.out_binding = -1
    import ctypes
    L.check(lib.cpn_conv2d(ctypes.byref(op), L.ptr(xs), L.ptr(out), L.ptr(rs), L.ptr(blob), L.stream_ptr()), 'conv2d')
    return out.permute(0, 3, 1, 2).float()
