"""Digest the artefacts a GPU run left in gpurun_out/ into small, committed summaries under profiles/ (round tag r01)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r01'


def launches_summary():
    path = os.path.join(OUT, 'launches.csv')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    tot, cnt = collections.Counter(), collections.Counter()
    for row in r:
        if len(row) <= vi:
            continue
        name = re.sub(r'\(.*', '', row[ki]).replace('void ', '').replace('cpn::', '')
        v = float(row[vi].replace(',', ''))
        v = v / 1e3 if row[ui] == 'ns' else (v * 1e3 if row[ui] == 'ms' else v)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    out = [f'# ncu launch list of the timed region of `bench.py --steps 2` (gpu__time_duration.sum, --clock-control none;',
           f'# serialised + cold-cache: compare SHARES).  {sum(cnt.values())} launches, {T / 1e3:.2f} ms total.', '']
    for k, v in tot.most_common(40):
        out.append(f'{v:12.1f} us {100 * v / T:6.2f}% {cnt[k]:6d}x  {k[:110]}')
    return '\n'.join(out) + '\n'


def ncu_raw(rep, keys):
    p = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True)
    rows = list(csv.reader(p.stdout.splitlines()))
    if len(rows) < 3:
        return None
    hdr, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h in keys or h in ('Kernel Name', 'Grid Size', 'Block Size'):
                d[h] = (row[i], units[i])
        res.append(d)
    return res


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum', 'smsp__inst_executed.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']


def main():
    os.makedirs(PROF, exist_ok=True)
    s = launches_summary()
    if s:
        open(os.path.join(PROF, f'{TAG}_launches_summary.txt'), 'w').write(s)
    for rep, name in (('prof_heads_conv.ncu-rep', 'ncu_heads_conv'), ('prof_f2c.ncu-rep', 'ncu_f2c')):
        path = os.path.join(OUT, rep)
        if os.path.exists(path):
            res = ncu_raw(path, KEYS)
            if res:
                with open(os.path.join(PROF, f'{TAG}_{name}.json'), 'w') as f:
                    json.dump(res, f, indent=1)
    for src, dst in (('plan_profile.txt', f'{TAG}_plan_profile.txt'), ('parity_report.json', f'{TAG}_parity_report.json'),
                     ('bench_fp16.log', f'{TAG}_bench_fp16.log'), ('bench_decode_minb3.log', f'{TAG}_bench_decode.log'),
                     ('bench_decode.log', f'{TAG}_bench_decode.log'), ('pytest.log', f'{TAG}_pytest_gpu.log')):
        p = os.path.join(OUT, src)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(PROF, dst))
    print(os.listdir(PROF))


if __name__ == '__main__':
    main()
