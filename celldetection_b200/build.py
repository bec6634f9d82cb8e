"""In-tree build of libcpn_b200.so with nvcc for sm_100a (no torch headers, no pybind: plain C ABI).

Usage: ``python -m celldetection_b200.build [--force] [--verbose]``; also called by ``__graft_entry__.build()``.
The shared object lands next to this file so it travels with the repository snapshot to the GPU box.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libcpn_b200.so')
STAMP = os.path.join(HERE, 'csrc', '.build_stamp')

# (source, extra flags).  post.cu / nms.cu must round like the reference's un-fused torch ops -> no FMA contraction.
SOURCES = [
    ('plan.cu', []),
    ('conv_simt.cu', []),
    ('conv_tc.cu', []),
    ('pointwise.cu', []),
    ('post.cu', ['-fmad=false']),
    ('nms.cu', ['-fmad=false']),
    ('labels.cu', ['-fmad=false']),   # shares the grid-cell arithmetic with nms.cu: must round identically
]
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
          '-Xcudafe', '--diag_suppress=177']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for name in sorted(os.listdir(root)):
            if name.endswith(('.cu', '.cuh', '.h')):
                with open(os.path.join(root, name), 'rb') as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(repr((SOURCES, ARCH, COMMON)).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src, extra in SOURCES:
        obj = os.path.join(CSRC, src.replace('.cu', '.o'))
        cmd = [nvcc] + ARCH + COMMON + extra + (['-Xptxas', '-v'] if verbose else []) + \
              ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f'--- nvcc failed for {src}\n{" ".join(cmd)}\n{out}\n')
        elif verbose or out.strip():
            sys.stderr.write(f'--- {src}\n{out}\n')
    if failed:
        raise RuntimeError('nvcc compilation failed')
    cmd = [nvcc] + ARCH + ['-shared', '-o', LIB] + objs + ['-lcudart']
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as f:
        f.write(digest)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
