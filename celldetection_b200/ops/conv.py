"""Stand-alone access to the plan's convolution kernels (per-layer parity tests and micro-benchmarks).

``conv2d`` takes/returns NCHW tensors like ``F.conv2d`` but runs ``cpn_conv2d``: NHWC staging, BatchNorm-free folded
weights, optional residual and ReLU in the epilogue.  ``engine='tcgen05'`` needs fp16-representable operands
(C_in per slab and C_out multiples of 64); ``engine='simt'`` is the strict fp32 CUDA-core kernel.
"""
import torch

from .. import _lib as L
from ..models.plan import expand_grouped, slab_of, pack_f16f8

__all__ = ['conv2d']


def _split_nhwc(t_nchw, c_pad):
    """NCHW fp32 -> NHWC (hi | lo) fp16 pairs with c_pad channels per half."""
    n, c, h, w = t_nchw.shape
    v = t_nchw.permute(0, 2, 3, 1).float()
    hi = v.half()
    lo = (v - hi.float()).half()
    out = torch.zeros((n, h, w, 2 * c_pad), dtype=torch.float16, device=t_nchw.device)
    out[..., :c] = hi
    out[..., c_pad:c_pad + c] = lo
    return out


def _f16f8_nhwc(t_nchw, c_pad):
    """NCHW fp32 -> NHWC CPN_DT_F16F8 pixels (cpn_b200.h): c_pad fp16 values, then per 32-channel chunk 32 bytes
    e4m3((v - hi) * 2^8) and 32 bytes e4m3(hi * 2^-2); returned as a fp16-typed tensor [n, h, w, 2 * c_pad]."""
    n, c, h, w = t_nchw.shape
    assert c_pad % 32 == 0
    v = torch.zeros((n, h, w, c_pad), dtype=torch.float32, device=t_nchw.device)
    v[..., :c] = t_nchw.permute(0, 2, 3, 1).float()
    hi = v.half()
    lo8 = ((v - hi.float()) * 2. ** 8).clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)
    hi8 = (hi.float() * 2. ** -2).clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)
    b8 = torch.stack((lo8.reshape(n, h, w, c_pad // 32, 32), hi8.reshape(n, h, w, c_pad // 32, 32)), 4)
    return torch.cat((hi, b8.reshape(n, h, w, 2 * c_pad).view(torch.float16)), 3).contiguous()


def _f16f8_decode(out, c, c_pad):
    """[n, h, w, 2 * c_pad] fp16-typed CPN_DT_F16F8 pixels -> (value = hi + lo8 * 2^-8, hi8 * 4) as NHWC fp32."""
    n, h, w, _ = out.shape
    b8 = out[..., c_pad:].contiguous().view(torch.uint8).reshape(n, h, w, c_pad // 32, 2, 32)
    lo8 = b8[..., 0, :].reshape(n, h, w, c_pad).contiguous().view(torch.float8_e4m3fn).float()
    hi8 = b8[..., 1, :].reshape(n, h, w, c_pad).contiguous().view(torch.float8_e4m3fn).float()
    return (out[..., :c].float() + lo8[..., :c] * 2. ** -8), hi8[..., :c] * 4.


def conv2d(x, weight, bias=None, stride=1, padding=0, groups=1, residual=None, relu=False, engine='simt',
           half=None):
    """x [N,Cin,H,W], weight [Cout,Cin/groups,k,k] (CUDA) -> [N,Cout,Ho,Wo] fp32.
    ``engine``: 'simt' | 'tcgen05' | 'tcgen05x3' (split fp16 pairs, three tensor-core passes) | 'tcgen05f8' (fp16 +
    e4m3 corrections, one fp16 and one fp8 tensor-core pass)."""
    if engine in ('tcgen05x3', 'tcgen05f8'):
        return _conv2d_split(x, weight, bias, stride, padding, groups, residual, relu, f8=engine == 'tcgen05f8')
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


def _conv2d_split(x, weight, bias, stride, padding, groups, residual, relu, f8=False):
    import ctypes
    lib = L.load()
    n, cin, h, w = x.shape
    cout, _, k, _ = weight.shape
    ho, wo = (h + 2 * padding - k) // stride + 1, (w + 2 * padding - k) // stride + 1
    al = 32 if f8 else 8
    cin_p, cout_p = (cin + al - 1) // al * al, (cout + al - 1) // al * al
    dt = L.DT_F16F8 if f8 else L.DT_F16X2
    stage = _f16f8_nhwc if f8 else _split_nhwc
    xs = stage(x, cin_p)
    out = torch.zeros((n, ho, wo, 2 * cout_p), dtype=torch.float16, device=x.device)
    wf = weight.float()
    if groups > 1:
        wf = expand_grouped(wf, groups)
    kslab, mode = slab_of(cin, cout, groups)
    wp = wf.permute(2, 3, 0, 1).reshape(k * k, cout, kslab).contiguous()
    acc_scale = 1.
    if f8:
        wb, acc_scale = pack_f16f8(wp.cpu())
        wb = wb.reshape(-1).to(x.device)
    else:
        hi = wp.half()
        lo = (wp - hi.float()).half()
        wb = torch.cat((hi, hi, lo), 2).contiguous().view(torch.uint8).reshape(-1)
    b_off = (wb.numel() + 255) // 256 * 256
    bb = (bias if bias is not None else torch.zeros(cout, device=x.device)).float().contiguous().view(torch.uint8)
    blob = torch.zeros(b_off + bb.numel() + 256, dtype=torch.uint8, device=x.device)
    blob[:wb.numel()] = wb
    blob[b_off:b_off + bb.numel()] = bb.reshape(-1)
    op = L.Op()
    op.kind, op.engine = L.OP_CONV, L.ENGINE_TCGEN05
    op.acc_scale = acc_scale
    for v, (c, hh, ww, cp) in ((op.src, (cin, h, w, cin_p)), (op.dst, (cout, ho, wo, cout_p))):
        v.offset, v.n, v.h, v.w, v.c, v.pitch, v.dtype, v.lo_delta = 0, n, hh, ww, c, 2 * cp, dt, cp
    rs = None
    if residual is not None:
        rs = stage(residual, cout_p)
        v = op.res
        v.offset, v.n, v.h, v.w, v.c, v.pitch, v.dtype, v.lo_delta = (0, n, residual.shape[2], residual.shape[3], cout,
                                                                      2 * cout_p, dt, cout_p)
    op.w_offset, op.b_offset = 0, b_off
    op.r = op.s = k
    op.stride, op.pad, op.kslab, op.slab_mode = stride, padding, kslab, mode
    op.act = L.ACT_RELU if relu else L.ACT_NONE
    op.out_binding = -1
    L.check(lib.cpn_conv2d(ctypes.byref(op), L.ptr(xs), L.ptr(out), L.ptr(rs), L.ptr(blob), L.stream_ptr()), 'conv2d')
    if f8:
        o, hi8 = _f16f8_decode(out, cout, cout_p)
        conv2d.last_hi8 = hi8.permute(0, 3, 1, 2).contiguous()     # for the tests: the e4m3 copy of the output
    else:
        o = out[..., :cout].float() + out[..., cout_p:cout_p + cout].float()
    return o.permute(0, 3, 1, 2).contiguous()
