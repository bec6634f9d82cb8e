"""Micro-bench of the slide-side integer kernels (cpn_histogram / cpn_apply_lut / cpn_rgb2gray / cpn_label_props): achieved
HBM GB/s (algorithmic bytes = input read + output written) against the measured copy peak.  CUDA events, 3 warm-ups."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from celldetection_b200 import _lib as L  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    lib = L.load()
    peak = 6550.
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peak = float(json.load(f).get('hbm_gbs', peak))
    except Exception:
        pass
    st = L.stream_ptr()
    g = torch.Generator(device='cuda').manual_seed(0)
    H = W = 16384
    rows = []
    # a microscopy-like distribution (concentrated low values) and a uniform one
    for name, make in (('gamma', lambda n, top: (torch.empty(n, device='cuda').exponential_(8., generator=g) * top).clamp_(0, top)),
                       ('uniform', lambda n, top: torch.rand(n, device='cuda', generator=g) * top)):
        u8 = make(H * W * 3, 255).to(torch.uint8)
        u16 = make(H * W, 65535).to(torch.int32).to(torch.int16)
        hist8 = torch.empty(256, dtype=torch.int32, device='cuda')
        hist16 = torch.empty(65536, dtype=torch.int32, device='cuda')
        lut8 = torch.randint(0, 256, (256,), dtype=torch.uint8, device='cuda')
        lut16 = torch.randint(0, 256, (65536,), dtype=torch.uint8, device='cuda')
        out8 = torch.empty_like(u8)
        out16 = torch.empty(u16.numel(), dtype=torch.uint8, device='cuda')
        gray = torch.empty(H * W, dtype=torch.uint8, device='cuda')
        cases = [
            ('histogram u8 ' + name, u8.numel(), lambda: lib.cpn_histogram(L.ptr(u8), L.DT_U8, u8.numel(), L.ptr(hist8), st)),
            ('histogram u16 ' + name, 2 * u16.numel(), lambda: lib.cpn_histogram(L.ptr(u16), L.DT_U16, u16.numel(), L.ptr(hist16), st)),
            ('apply_lut u8 ' + name, 2 * u8.numel(), lambda: lib.cpn_apply_lut(L.ptr(u8), L.DT_U8, u8.numel(), L.ptr(lut8), L.ptr(out8), st)),
            ('apply_lut u16 ' + name, 3 * u16.numel(), lambda: lib.cpn_apply_lut(L.ptr(u16), L.DT_U16, u16.numel(), L.ptr(lut16), L.ptr(out16), st)),
            ('rgb2gray u8 ' + name, 4 * H * W, lambda: lib.cpn_rgb2gray(L.ptr(u8), L.DT_U8, H * W, 3, None, L.ptr(gray), st)),
        ]
        for label, nbytes, fn in cases:
            assert fn() == 0, L.load().cpn_last_error()
            ms = timed(fn)
            rows.append((label, nbytes, ms))
        assert int(hist8.sum()) == u8.numel() and int(hist16.to(torch.int64).sum()) == u16.numel()
        del u8, u16, out8, out16, gray
    # label statistics: a flat label image with ~70 000 disc-like regions on 16384^2 (and a 4-channel one at 8192^2)
    for (h, w, c) in ((16384, 16384, 1), (8192, 8192, 4)):
        yy = torch.arange(h, device='cuda', dtype=torch.int32)[:, None]
        xx = torch.arange(w, device='cuda', dtype=torch.int32)[None, :]
        cell = (yy // 64) * (w // 64) + (xx // 64) + 1
        inside = ((yy % 64 - 32) ** 2 + (xx % 64 - 32) ** 2) < 20 ** 2
        lab = torch.where(inside, cell, torch.zeros_like(cell))
        lab = lab[..., None].expand(h, w, c).contiguous()
        mx = int(lab.max())
        slots = c * (mx + 1)
        area = torch.empty(slots, dtype=torch.int32, device='cuda')
        bbox = torch.empty((slots, 4), dtype=torch.int32, device='cuda')
        sums = torch.empty((slots, 2), dtype=torch.int64, device='cuda')
        flags = torch.empty(1, dtype=torch.int32, device='cuda')
        fn = lambda: lib.cpn_label_props(L.ptr(lab), h, w, c, mx, L.ptr(area), L.ptr(bbox), L.ptr(sums), L.ptr(flags), st)  # noqa: E731
        assert fn() == 0, L.load().cpn_last_error()
        ms = timed(fn)
        assert int(area.to(torch.int64).sum()) == int((lab > 0).sum())
        rows.append((f'label_props {h}x{w}x{c} ({mx} regions)', lab.numel() * 4, ms))
        del lab
    for label, nbytes, ms in rows:
        gbps = nbytes / ms / 1e6
        print(f'{label:44s} {nbytes / 1e6:9.1f} MB  {ms:8.4f} ms  {gbps:8.1f} GB/s  {gbps / peak:5.2f} of {peak:.0f}')


if __name__ == '__main__':
    main()
