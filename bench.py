#!/usr/bin/env python
"""Headline benchmark: tiles/sec (3x512x512) of CpnResNeXt101UNet inference on N B200s of one node, in the engine that
meets north_star's parity gate (``fp16f8``: fp16 tensor-core pass + one e4m3 correction pass, head tensors within 1e-3 of
the reference, identical instance counts).

  python bench.py [--gpus 1] [--steps K] [--warmup W]          # N = 1: BASELINE.json configs[2] (C3), batch 16x3x512x512
  python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...
                                                                # N > 1: configs[3] (C4), ONE 3x16384x16384 slide, 1849
                                                                # tiles sharded over the ranks + NCCL all-gather + stitch
  python bench.py --impl reference ...                          # the reference's CPU path (oracle port) on host cores

N = 1: a "step" is one full pass of the hot path over one batch (backbone + heads + select / decode / refine / boxes +
NMS); ``value`` times it with the batch resident in HBM, ``e2e`` times ``model(x)`` from pinned HOST memory with the
results copied back to the host inside the timed region.  N > 1: a "step" is one whole slide through ``cd.apply_model``
(tile loop on every rank, border filter, all-gather of the detection records, canonical re-sort, global stitch NMS);
``value`` with the slide resident in HBM, ``e2e`` from the host image (pinned double-buffered staging) with the stitched
result copied back; strong scaling (the work is fixed), the per-rank replica throughput is reported beside it.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH = 'CpnResNeXt101UNet'
TILE = 512
BATCH = 16
SEED = 0
FG_FRACTION = 0.02     # SURVEY 8(d) calibration: 2 % of head pixels above score_thresh
METRIC = 'tiles/sec (3x512x512) CpnResNeXt101UNet inference'
HEADLINE = 'fp16f8'    # default engine: the tensor-core engine that meets the 1e-3 parity gate
WSI, CROP, STRIDE = 16384, 512, 384   # C4: 43 x 43 = 1849 tiles
DTYPE = {'fp16f8': 'f16+e4m3', 'fp16': 'f16', 'fp16x3': 'f16x3', 'fp32': 'f32'}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d['hbm_gbs'], tflops=d['bf16_tflops'], tflops_sustained=d['bf16_tflops_sustained'],
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650., tflops=1590., tflops_sustained=1400., source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.p = index, None

    def __enter__(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                       '-lms', '200', '-i', str(self.index)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None
        return self

    def __exit__(self, *a):
        self.result = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            return
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [c.strip() for c in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if sm:
            sm.sort()
            self.result = dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def build_state_dict(core_fn, calib_x, arch=None):
    from celldetection_b200.models.graph import trace
    from celldetection_b200.utils.synth import synth_state_dict, calibrate_heads_
    spec = trace(arch or ARCH, 1, 64, 64).spec
    sd = synth_state_dict(spec, seed=SEED)
    calibrate_heads_(sd, core_fn, calib_x, fg_fraction=FG_FRACTION, fourier_std=3.0, location_std=1.0)
    return sd


def synthetic_batches(n_batches, device, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(BATCH, 3, TILE, TILE, generator=g).to(device) for _ in range(n_batches)]


# ----------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU path (oracle port, torch CPU fp32 on all host threads)
# ----------------------------------------------------------------------------------------------------------------------
def host_threads():
    """Usable host threads: min(os.cpu_count(), scheduler affinity, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:
            q, p = f.read().split()
            if q != 'max':
                n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    return max(1, n)


def cpu_reference_tiles_per_sec(sd, steps, warmup, tiles_per_step=1):
    """Times the oracle port (torch-CPU fp32) of the whole path on 1-tile steps.  The thread count is chosen among
    {T, T/2, T/4} (T = usable host threads) by timing one warm-up tile each: oneDNN does not always scale to every
    hardware thread on a batch of one, and the baseline should be the best the host can do."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import cpn_oracle as orc
    total = host_threads()
    g = torch.Generator().manual_seed(SEED + 17)
    xs = [torch.rand(tiles_per_step, 3, TILE, TILE, generator=g) for _ in range(2)]
    best, cores = None, total
    for cand in sorted({total, max(1, total // 2), max(1, total // 4)}, reverse=True):
        torch.set_num_threads(cand)
        if best is None:
            orc.cpn_forward(xs[0], sd, ARCH)              # first-touch warm-up (allocator, oneDNN primitives)
        t0 = time.perf_counter()
        orc.cpn_forward(xs[1], sd, ARCH)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, cores = dt, cand
    torch.set_num_threads(cores)
    kept = 0
    for i in range(max(0, warmup - 1)):
        orc.cpn_forward(xs[i % 2], sd, ARCH)
    t0 = time.perf_counter()
    for i in range(steps):
        out = orc.cpn_forward(xs[i % 2], sd, ARCH)
        kept = sum(len(s) for s in out['scores'])
    dt = time.perf_counter() - t0
    return tiles_per_step * steps / dt, dt / steps * 1e3, cores, kept


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path (oracle port: torch-CPU fp32 + numpy) on the
    box's host cores.  N = 1: C3 tiles; N > 1: the C4 tiled driver (``apply_model``: tiles, border removal, global NMS)
    on a bounded 1280 x 1280 corner of the same synthetic slide (3 x 3 tiles at crop 512 / stride 384)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import cpn_oracle as orc

    def core_fn(x, sd_):
        with torch.no_grad():
            s, l, r, f = orc.cpn_core(x, sd_, ARCH)
        return dict(scores=s, locations=l, fourier=f, refinement=r)

    torch.set_num_threads(host_threads())
    g = torch.Generator().manual_seed(SEED)
    calib = torch.rand(1, 3, TILE, TILE, generator=g)
    sd = build_state_dict(core_fn, calib)
    steps_in = args.steps if args.steps is not None else 3
    if args.gpus > 1 and args.workload != 'c3':
        import numpy as np
        steps, warmup = max(1, min(steps_in, 2)), 1
        side = 2 * STRIDE + CROP                                   # 1280: 3 x 3 tiles
        rng = np.random.RandomState(SEED)
        img = rng.randint(0, 256, size=(side, side, 3), dtype=np.uint8)
        _, cores = _pick_threads(lambda: orc.cpn_forward(torch.rand(1, 3, TILE, TILE), sd, ARCH))
        with torch.no_grad():
            for _ in range(warmup - 1):
                orc.apply_model(img, sd, ARCH, CROP, STRIDE)
            t0 = time.perf_counter()
            for _ in range(steps):
                res = orc.apply_model(img, sd, ARCH, CROP, STRIDE)
            dt = time.perf_counter() - t0
        ntiles = 9
        val, ms, kept = ntiles * steps / dt, dt / steps * 1e3, int(res['scores'].shape[0])
        workload = (f'C4 sample: {ARCH} random-init (calibrated heads), {side}x{side} corner of the synthetic slide, crop '
                    f'{CROP} / stride {STRIDE} -> {ntiles} tiles, border removal + global NMS, on CPU')
        sample = f'{steps} steps x {ntiles} tiles (one {side}x{side} image through oracle apply_model), torch-CPU fp32'
        scaling = 'strong'
    else:
        steps, warmup = max(1, min(steps_in, 8)), max(1, min(args.warmup, 2))
        val, ms, cores, kept = cpu_reference_tiles_per_sec(sd, steps, warmup)
        workload = f'C3: {ARCH} random-init (calibrated heads), 1x3x{TILE}x{TILE} per step on CPU'
        sample = f'{steps} steps x 1 tile of 3x{TILE}x{TILE}, oracle/cpn_oracle.py torch-CPU fp32'
        scaling = 'weak'
    line = dict(impl='reference', metric=METRIC, value=val, unit='tiles/s', n_gpus=args.gpus, steps=steps,
                warmup=warmup, ms_per_step=ms, higher_is_better=True, scaling=scaling, vs_baseline=None, dtype='f32',
                data='synthetic', config=dict(workload=workload, tile=TILE, kept_last_step=kept),
                cpu_baseline=dict(value=val, unit='tiles/s', cores=cores, kind='port', sample=sample),
                e2e=dict(value=val, unit='tiles/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def _pick_threads(fn):
    """Best of {T, T/2, T/4} host threads for `fn` (one call each after a warm-up).  Sets and returns (seconds, threads)."""
    total = host_threads()
    best, cores = None, total
    with torch.no_grad():
        for cand in sorted({total, max(1, total // 2), max(1, total // 4)}, reverse=True):
            torch.set_num_threads(cand)
            if best is None:
                fn()
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, cores = dt, cand
    torch.set_num_threads(cores)
    return best, cores


# ----------------------------------------------------------------------------------------------------------------------
# this repository's arm
# ----------------------------------------------------------------------------------------------------------------------
def make_model(arch, precision, dev, seed_offset=0):
    """Random-init model of `arch` with calibrated heads, calibrated through its own engine on `dev`."""
    import celldetection_b200 as cd
    model = getattr(cd.models, arch)(3, precision=precision)
    g = torch.Generator().manual_seed(SEED + seed_offset)
    calib = torch.rand(1, 3, TILE, TILE, generator=g)

    def core_fn(x, sd_):
        model.load_state_dict(sd_)
        model.to(dev)
        out = model.core_forward(x.to(dev))
        return {k: v.float().cpu() for k, v in out.items()}

    sd = build_state_dict(core_fn, calib, arch)
    model.load_state_dict(sd)
    model.to(dev)
    return model, sd


def sibling(arch, precision, sd, dev):
    import celldetection_b200 as cd
    m = getattr(cd.models, arch)(3, precision=precision)
    m.load_state_dict(sd)
    return m.to(dev)


def time_steps(fn, steps, warmup, barrier, dist=None, dev=None, profile_name=None):
    """W untimed warm-up steps, then exactly K steps between CUDA events, bracketed by barrier + synchronize; max over
    ranks.  Returns (ms total, library launches inside the timed region, last result)."""
    from celldetection_b200 import _lib
    last = None
    for i in range(warmup):
        last = fn(i)
    barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = profile_name is not None and os.environ.get('CPN_PROFILE_RANGE') == profile_name   # ncu --profile-from-start off
    if prof:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(steps):
        last = fn(i)
    e1.record()
    barrier()
    if prof:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, _lib.launch_count() - l0, last


def heads_roofline(model, x, precision):
    """Dominant kernel, timed alone with CUDA events: the merged 7x7 head convolution (conv_halo_kernel<256,1>) with the
    fused ReadOut projections.  achieved = ALGORITHMIC FLOPs (2 * N * h * w * C_out * C_in * 49, what the reference's
    conv computes) / launch time; the 2-pass engine executes twice that many tensor-core pass-equivalents."""
    from celldetection_b200 import _lib
    peaks = measured_peaks()
    plan = model._plan(BATCH, TILE, TILE)
    head_idx = [i for i, o in enumerate(plan.g.ops) if o.name == 'heads.block.0'][0]
    hop = plan.g.ops[head_idx]
    outs = plan.new_outputs()
    plan.forward(x, _lib.IN_F32_NCHW, outs)               # fills the head features the op reads
    for _ in range(3):
        plan.run_op(head_idx, x, _lib.IN_F32_NCHW, outs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        plan.run_op(head_idx, x, _lib.IN_F32_NCHW, outs)
    e1.record()
    torch.cuda.synchronize()
    hms = e0.elapsed_time(e1) / reps
    hflops = 2. * BATCH * hop.dst.h * hop.dst.w * hop.dst.c * hop.src.c * hop.k * hop.k
    achieved = hflops / (hms / 1e3) / 1e12
    passes = {'fp16f8': 2, 'fp16': 1, 'fp16x3': 3}.get(precision)
    if passes is None:
        return dict(bound='tensor', kernel='conv_simt_kernel<float> (strict fp32 CUDA-core engine)', achieved=achieved,
                    peak=peaks['tflops'], unit='TFLOP/s', frac=achieved / peaks['tflops'], traffic=None,
                    peak_source=peaks['source'], ms_per_launch=hms, flops_per_launch=hflops)
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full` capture of the same
    # launch (a profiler cannot run inside the bench); null until a capture of this engine is committed
    traffic, src = None, None
    for cand in ('r02_final_ncu_c3_heads.json' if precision == 'fp16f8' else None,      # final-state capture of this kernel
                 f'r02_ncu_heads_conv_{precision}.json', 'r01_ncu_heads_conv.json' if precision == 'fp16' else None):
        path = cand and os.path.join(ROOT, 'profiles', cand)
        if path and os.path.exists(path):
            try:
                with open(path) as f:
                    r0 = json.load(f)[0]
                unit = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
                traffic = sum(float(r0[k][0]) * unit[r0[k][1]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
                src = 'profiles/' + cand
                break
            except Exception:
                traffic = None
    act_bytes = {'fp16f8': 4, 'fp16': 2, 'fp16x3': 4}[precision]      # bytes per activation / weight element as stored
    return dict(bound='tensor',
                kernel=f'conv_halo_kernel<256,1>: 7x7 head convolution {hop.src.c}->{hop.dst.c} @{hop.dst.h}x{hop.dst.w} + fused '
                       f'ReadOut projection(s)' + (' (score head; location / fourier heads are evaluated at proposals only)'
                                                   if getattr(plan.g, 'sparse', False) else ' (merged score / location / fourier heads)')
                       + f', batch {BATCH}, {passes} tcgen05 pass-equivalent(s) per K block'
                       + (' (kind::f8f6f4 e4m3 correction pass, then kind::f16; one fp32 accumulator)' if precision == 'fp16f8' else ''),
                achieved=achieved, peak=peaks['tflops'], unit='TFLOP/s', frac=achieved / peaks['tflops'],
                traffic=traffic, traffic_source=src,
                algorithmic_bytes=BATCH * hop.src.h * hop.src.w * hop.src.c * act_bytes
                + hop.dst.c * hop.src.c * hop.k * hop.k * act_bytes
                + BATCH * hop.dst.h * hop.dst.w * (1 if getattr(plan.g, 'sparse', False) else 23) * 4,
                peak_source=peaks['source'] + ', burst cuBLAS bf16', ms_per_launch=hms, flops_per_launch=hflops,
                executed_tflops_pass_equivalents=achieved * passes,
                executed_frac_of_peak=achieved * passes / peaks['tflops'])


def measured_parity(precision, dev):
    """Parity of the benchmarked engine, MEASURED in this run on the reference-minted golden vector of the flagship
    architecture (tests/golden/model_cpnresnext101unet_n1_128.npz: raw head tensors and outputs of the unmodified
    reference on the same seeded input and state_dict)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    try:
        from helpers import load_npz, fixture_state_dict, rel_err, match_by_box
        import celldetection_b200 as cd
        import numpy as np
        name = 'model_cpnresnext101unet_n1_128'
        z = load_npz(name)
        arch = str(z['arch'])
        n, h, w, seed, order, samples = [int(v) for v in z['meta']]
        m = getattr(cd.models, arch)(3, order=order, samples=samples, precision=precision)
        m.load_state_dict(fixture_state_dict(z, arch, seed))
        m = m.to(dev)
        x = torch.from_numpy(z['x']).to(dev)
        raw = m.core_forward(x)
        errs = {k: rel_err(raw[k].cpu().numpy(), z['raw_' + k]) for k in ('scores', 'locations', 'refinement', 'fourier')}
        out = m(x)
        rb = z['out/0/boxes']
        pairs = match_by_box(out['boxes'][0].cpu().numpy(), rb)
        verr = max([float(np.abs(out['contour_proposals'][0][a].cpu().numpy() - z['out/0/contour_proposals'][b]).max())
                    for a, b in pairs] or [0.])
        gate = max(errs.values()) < 1e-3 and len(out['scores'][0]) == len(rb) and verr < 0.5
        return dict(gate='head tensors within 1e-3 rel (max|a-b| / max|b|), decoded contour vertices within 0.5 px, '
                         'identical instance count after NMS', fixture=name, precision=precision,
                    head_tensor_rel_err=errs, instances=len(out['scores'][0]), reference_instances=int(len(rb)),
                    matched=len(pairs), decoded_vertex_err_px=verr, meets_gate=bool(gate), measured_in_this_run=True)
    except Exception as e:   # the goldens are part of the repository; report rather than hide a failure
        return dict(error=f'{type(e).__name__}: {e}'[:300], measured_in_this_run=False)


def quick_rate(model, xs, steps=5, warmup=2):
    for i in range(warmup):
        model.forward_flat(xs[i % len(xs)])
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for i in range(steps):
        flat, counts = model.forward_flat(xs[i % len(xs)])
    a1.record()
    torch.cuda.synchronize()
    ms = a0.elapsed_time(a1) / steps
    return ms, int(sum(counts))


def batch1_latency(model, x1):
    """Device-synchronised latency of ``model(x)`` for ONE 3x512x512 tile (median of 20), launched kernel by kernel and as a
    CUDA-graph replay of the plan."""
    out = {}
    for name, flag in (('eager_ms', False), ('cuda_graph_ms', True)):
        model.cuda_graph = flag
        for _ in range(3):
            model(x1)
        ts = []
        for _ in range(20):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            model(x1)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        out[name] = ts[len(ts) // 2]
    model.cuda_graph = False
    out['note'] = 'one tile, host wall clock around model(x) incl. the post-head chain and its host read-backs'
    return out


def config_c2(dev, peaks):
    """BASELINE configs[1]: CpnResNet18FPN, batch 32x3x512x512, headline engine (side leg, a few steps)."""
    from celldetection_b200.models.graph import conv_flops
    arch = 'CpnResNet18FPN'
    m, _ = make_model(arch, HEADLINE, dev, seed_offset=5)
    g = torch.Generator().manual_seed(SEED + 31)
    xs = [torch.rand(32, 3, TILE, TILE, generator=g).to(dev) for _ in range(2)]
    ms, kept = quick_rate(m, xs, steps=5, warmup=2)
    fl = conv_flops(m._plan(32, TILE, TILE).g)
    tf = fl / (ms / 1e3) / 1e12
    return dict(workload=f'{arch}, batch 32x3x{TILE}x{TILE}, {HEADLINE}', value=32 / (ms / 1e3), unit='tiles/s',
                ms_per_step=ms, kept_last_step=kept, conv_tflops=tf, frac_of_sustained_peak=tf / peaks['tflops_sustained'],
                note='conv_tflops counts the reference formulation (7x7 refinement head on the bilinearly up-sampled 512^2 '
                     'features); the plan runs that head as four 5x5 phase convolutions on the 256^2 features + border strips')


def config_c5(dev, peaks):
    """BASELINE configs[4]: fouriers2contours, 1e6 proposals x order 16 x 128 samples (1288 algorithmic bytes each)."""
    import celldetection_b200 as cd
    from celldetection_b200 import _lib as L
    P, ORDER, S = 1_000_000, 16, 128
    lib = L.load()
    g = torch.Generator(device=dev).manual_seed(0)
    sets = [(torch.randn(P, ORDER, 4, device=dev, generator=g), torch.rand(P, 2, device=dev, generator=g) * 512,
             torch.empty(P, S, 2, device=dev)) for _ in range(4)]      # 4 x 1.29 GB: nothing re-used out of L2
    trig = cd.ops.cpn.trig_table(ORDER, S, dev)

    def run(i):
        f, l, o = sets[i % 4]
        L.check(lib.cpn_fouriers2contours(L.ptr(f), L.ptr(l), P, ORDER, S, L.ptr(trig), None, L.ptr(o), L.stream_ptr()))
    for i in range(4):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byt = P * (16 * ORDER + 8 + 8 * S)
    gbs = byt / (ms / 1e3) / 1e9
    return dict(workload='fouriers2contours 1e6 proposals x order 16 x 128 samples', value=P / (ms / 1e3) / 1e6,
                unit='Mproposals/s', ms_per_launch=ms,
                roofline=dict(bound='hbm', kernel='f2c_fast_kernel', achieved=gbs, peak=peaks['hbm_gbs'], unit='GB/s',
                              frac=gbs / peaks['hbm_gbs'], algorithmic_bytes=byt,
                              traffic=1.229e9, traffic_source='profiles/r01_ncu_f2c.json (264 MB read + 965 MB written)'))


def digest_of(res):
    h = hashlib.sha256()
    for k in ('boxes', 'scores', 'contours'):
        h.update(res[k].cpu().numpy().tobytes())
    return h.hexdigest()[:16]


def run_b200(args):
    import celldetection_b200 as cd
    from celldetection_b200 import _lib
    from celldetection_b200.models.graph import conv_flops
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    _lib.load()
    peaks = measured_peaks()
    c5 = None
    if world == 1 and args.workload in ('auto', 'c3') and not args.quick:
        # the decode micro-benchmark is a kernel timed alone: run it before the long tensor-core loops pull the clocks
        # down to the power cap (measured: 0.29 ms on an idle GPU, 0.45 ms right after the C2 / C3 legs)
        try:
            c5 = config_c5(dev, peaks)
        except Exception as e:
            c5 = dict(error=f'{type(e).__name__}: {e}'[:300])
        torch.cuda.empty_cache()
    model, sd = make_model(ARCH, args.precision, dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n_rot = 4   # rotate 4 distinct batches (201 MB of inputs; per-step activations are several GB) -> no L2 reuse
    xs = synthetic_batches(n_rot, dev, SEED + 1 + rank)
    workload = args.workload if args.workload != 'auto' else ('c3' if world == 1 else 'c4')

    if workload == 'c3':
        host = [x.cpu().pin_memory() for x in xs]

        def step(i):
            return model.forward_flat(xs[i % n_rot])

        def step_e2e(i):
            x = host[i % n_rot].to(dev, non_blocking=True)    # host -> device copy of this step's inputs (pinned)
            y = model(x)                                      # the call a user makes (reference API: per-image lists)
            # device -> host read of the whole result: one copy per key (the per-image tensors are consecutive row
            # ranges), split again on the host
            sizes = [len(s) for s in y['scores']]
            out = {k: list(torch.split(torch.cat(v).cpu(), sizes)) for k, v in y.items() if v is not None}
            return out, sizes

        with ClockSampler(local) as clk:
            ms, launches, last = time_steps(step, args.steps, args.warmup, barrier, dist, dev, 'step')
        flat, counts = last
        ms_e2e, _, last_e = time_steps(step_e2e, args.steps, max(1, args.warmup), barrier, dist, dev)
        tiles = BATCH * world
        value = tiles * args.steps / (ms / 1e3)
        e2e = tiles * args.steps / (ms_e2e / 1e3)
        d2h = sum(t.numel() * t.element_size() for v in last_e[0].values() for t in v)
        h2d = int(host[0].numel() * 4)
        plan = model._plan(BATCH, TILE, TILE)
        from celldetection_b200.models.graph import trace as _trace
        algo_flops = conv_flops(_trace(ARCH, BATCH, TILE, TILE, stem_im2col=True))     # every head on every pixel
        total_flops = conv_flops(plan.g)                                                # what the plan executes per step
        sparse_rows = getattr(model, 'last_sparse_rows', 0) if getattr(plan.g, 'sparse', False) else 0
        if sparse_rows:                                                                 # + the heads at the proposals
            total_flops += 2. * sparse_rows * 2 * plan.g.head_mid * plan.g.head_feat.c * plan.g.head_k ** 2
        net_tflops = algo_flops * args.steps / (ms / 1e3) / 1e12
        exe_tflops = total_flops * args.steps / (ms / 1e3) / 1e12
        cfg = dict(workload=f'C3: {ARCH} random-init (synthetic weights seed {SEED}, heads calibrated to '
                            f'{FG_FRACTION:.0%} foreground), batch {BATCH}x3x{TILE}x{TILE} per GPU',
                   global_batch=tiles, tile=TILE, parallelism=f'tile-parallel x{world}',
                   l2='4 rotating input batches; per-step activations >> 126 MB L2',
                   proposals_last_step=int(sum(model.forward_flat(xs[0], nms=False)[1])),
                   kept_last_step=int(sum(counts)), precision=args.precision,
                   conv_gflop_per_tile=algo_flops / BATCH / 1e9, conv_gflop_per_tile_executed=total_flops / BATCH / 1e9,
                   sparse_heads=bool(getattr(plan.g, 'sparse', False)), sparse_head_rows_last_step=int(sparse_rows))
        line = dict(metric=METRIC, value=value, unit='tiles/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype=DTYPE[args.precision], data='synthetic', config=cfg, clocks=clk.result,
                    e2e=dict(value=e2e, unit='tiles/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=int(d2h),
                             ms_per_step=ms_e2e / args.steps, steps=args.steps),
                    gpu_launches=int(launches),
                    network=dict(conv_tflops=net_tflops, frac_of_sustained_peak=net_tflops / peaks['tflops_sustained'],
                                 frac_of_burst_peak=net_tflops / peaks['tflops'],
                                 conv_tflops_executed=exe_tflops,
                                 executed_frac_of_sustained_peak=exe_tflops / peaks['tflops_sustained'],
                                 launches_per_step=launches / args.steps,
                                 note='conv_tflops: ALGORITHMIC conv FLOPs (every head on every pixel, like the reference: '
                                      '2347.7 GF per tile after the commuted 1x1) / step time.  conv_tflops_executed: the '
                                      'FLOPs the plan really contracts (location / fourier heads only at the proposals), before '
                                      'the factor 2 of the 2-pass engine'))
    else:
        # ---- C4: one slide, tiles sharded over the ranks, all-gather + stitch on every rank ----
        import numpy as np
        size = args.wsi_size
        rng = np.random.RandomState(SEED)
        img = rng.randint(0, 256, size=(size, size, 3), dtype=np.uint8)          # identical on every rank
        img_dev = torch.from_numpy(img).to(dev)                                    # the slide resident in HBM
        _, shape = cd.get_tiling_slices((size, size), CROP, STRIDE)
        ntiles = shape[0] * shape[1]
        stages = {}

        def step(i):
            return cd.apply_model(img_dev, [model], crop_size=CROP, strides=STRIDE, batch_size=BATCH)

        def step_e2e(i):
            res = cd.apply_model(img, [model], crop_size=CROP, strides=STRIDE, batch_size=BATCH)
            return {k: v.cpu() for k, v in res.items()}      # the stitched result on the host

        with ClockSampler(local) as clk:
            ms, launches, res = time_steps(step, args.steps, args.warmup, barrier, dist, dev, 'step')
        e_steps = max(1, min(args.steps, 3))
        ms_e2e, _, res_e = time_steps(step_e2e, e_steps, 1, barrier, dist, dev)
        cd.apply_model(img_dev, [model], crop_size=CROP, strides=STRIDE, batch_size=BATCH, timings=stages)
        st = torch.tensor([stages['tiles_s'], stages['exchange_s'], stages['stitch_s']], device=dev)
        if dist is not None:
            dist.all_reduce(st, op=dist.ReduceOp.MAX)
        st = st.tolist()
        value = ntiles * args.steps / (ms / 1e3)
        e2e = ntiles * e_steps / (ms_e2e / 1e3)
        dg = digest_of(res)
        same = True
        if dist is not None:                               # every rank must hold the identical stitched result
            dgs = [None] * world
            dist.all_gather_object(dgs, dg)
            same = len(set(dgs)) == 1
        # replica throughput beside it: every rank its own C3 batch loop, no collective (round 1's weak-scaling number)
        rms, _ = quick_rate(model, xs, steps=3, warmup=1)
        rt = torch.tensor([rms], device=dev)
        if dist is not None:
            dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        rms = float(rt.item())
        mine = len(range(rank, ntiles, world))
        d2h = sum(v.numel() * v.element_size() for v in res_e.values())
        cfg = dict(workload=f'C4: {ARCH} random-init (calibrated heads), ONE 3x{size}x{size} uint8 synthetic slide, crop '
                            f'{CROP} / stride {STRIDE} -> {ntiles} tiles dealt round-robin to {world} GPUs (batch {BATCH}), '
                            f'border filter, NCCL all-gather of the detection records, canonical re-sort, global stitch NMS '
                            f'on every rank',
                   tiles=ntiles, tiles_per_rank=mine, tile=CROP, stride=STRIDE, parallelism=f'tile-parallel x{world}',
                   l2='every batch is new image data; per-batch activations >> 126 MB L2', precision=args.precision,
                   detections=int(res['scores'].shape[0]), digest=dg, identical_on_all_ranks=bool(same),
                   stage_seconds=dict(tile_loop=st[0], exchange=st[1], stitch_nms=st[2], note='max over ranks, one extra '
                                      'instrumented slide (device-synchronised between stages)'))
        line = dict(metric=METRIC, value=value, unit='tiles/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='strong', vs_baseline=None,
                    dtype=DTYPE[args.precision], data='synthetic', config=cfg, clocks=clk.result,
                    e2e=dict(value=e2e, unit='tiles/s', h2d_bytes_per_step=int(mine * CROP * CROP * 3),
                             d2h_bytes_per_step=int(d2h), ms_per_step=ms_e2e / e_steps, steps=e_steps,
                             note='host image -> pinned double-buffered crops -> device; stitched result copied back'),
                    gpu_launches=int(launches),
                    replicas=dict(value=BATCH * world / (rms / 1e3), unit='tiles/s', scaling='weak',
                                  note='every rank its own 16-tile batch loop (C3), no data-path collective'))

    if rank == 0:
        line['roofline'] = heads_roofline(model, xs[0], args.precision)
        line['parity'] = measured_parity(args.precision, dev)
        if world == 1 and workload == 'c3' and not args.quick:
            try:
                line['latency_batch1'] = batch1_latency(model, xs[0][:1].contiguous())
            except Exception as e:
                line['latency_batch1'] = dict(error=f'{type(e).__name__}: {e}'[:300])
            if getattr(model, 'sparse_heads', False):      # same engine, every head on every pixel (round-2 start)
                model.sparse_heads = False
                dms, dkept = quick_rate(model, xs)
                model.sparse_heads = True
                line['dense_heads'] = dict(value=BATCH / (dms / 1e3), unit='tiles/s', ms_per_step=dms, kept_last_step=dkept,
                                           note='location / fourier heads computed on every pixel like the reference does')
            if args.precision == HEADLINE:
                fast = sibling(ARCH, 'fp16', sd, dev)
                fms, fkept = quick_rate(fast, xs)
                line['fast_engine'] = dict(precision='fp16', value=BATCH / (fms / 1e3), unit='tiles/s', ms_per_step=fms,
                                           kept_last_step=fkept, meets_gate=False,
                                           note='single-pass fp16 tensor-core engine (opt-in): head tensors 1.3e-3..1e-2 '
                                                'from the reference, outside the 1e-3 gate')
                del fast
            try:
                line['configs'] = dict(C2=config_c2(dev, peaks), C5=c5)
            except Exception as e:
                line['configs'] = dict(error=f'{type(e).__name__}: {e}'[:300], C5=c5)
        if world == 1 and not args.no_cpu_baseline:
            sd_cpu = {k: v.cpu() for k, v in sd.items()}
            v, cms, cores, kept = cpu_reference_tiles_per_sec(sd_cpu, 3, 1)
            line['cpu_baseline'] = dict(value=v, unit='tiles/s', cores=cores, kind='port',
                                        sample=f'3 steps x 1 tile of 3x{TILE}x{TILE} (+1 warm-up), oracle/cpn_oracle.py '
                                               f'torch-CPU fp32')
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default=HEADLINE, choices=['fp16f8', 'fp16', 'fp16x3', 'fp32'])
    ap.add_argument('--workload', default='auto', choices=['auto', 'c3', 'c4'],
                    help='auto: C3 batch loop at N = 1, the sharded C4 slide at N > 1')
    ap.add_argument('--wsi-size', type=int, default=WSI)
    ap.add_argument('--quick', action='store_true', help='skip the side legs (fast engine, C2, C5)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.steps is None:      # C3: 20 batches (~1.4 s); C4: 5 whole slides
        args.steps = 20 if (args.workload == 'c3' or (args.workload == 'auto' and world == 1)) else 5
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
