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


KEYS = ['gpu__time_duration.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_uniform.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
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
    # one `ncu --set full` capture per kernel class (tools/gpu_r01b.sh): compact table of the roofline-relevant metrics
    for rep, name in (('prof_convs.ncu-rep', 'ncu_conv_classes'), ('prof_convs_f8.ncu-rep', 'ncu_conv_classes_fp16f8'),
                      ('prof_post.ncu-rep', 'ncu_post_chain')):
        path = os.path.join(OUT, rep)
        if not os.path.exists(path):
            continue
        res = ncu_raw(path, KEYS)
        if not res:
            continue
        with open(os.path.join(PROF, f'{TAG}_{name}.json'), 'w') as f:
            json.dump(res, f, indent=1)

        def val(d, k, scale=1.):
            try:
                v, u = d[k]
                v = float(v.replace(',', ''))
                if k == 'gpu__time_duration.sum':
                    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)        # -> us
                if k.endswith('bytes.sum') or k.startswith('dram__bytes') or k.startswith('launch__shared'):
                    v = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u.split('/')[0], 1)
                return v * scale
            except Exception:
                return float('nan')
        lines = ['# ncu --set full --clock-control none, one launch per kernel class (times are cold-cache, serialised)',
                 f'# {"kernel":58s} {"grid":>6s} {"us":>9s} {"tensor%":>8s} {"dram%":>6s} {"dramGB/s":>9s} {"readMB":>8s} '
                 f'{"writeMB":>8s} {"L2hit%":>7s} {"regs":>5s} {"smemKB":>7s}']
        for d in res:
            kn = re.sub(r'\(.*', '', d['Kernel Name'][0]).replace('void ', '').replace('cpn::', '')[:58]
            us = val(d, 'gpu__time_duration.sum')
            rd, wr = val(d, 'dram__bytes_read.sum'), val(d, 'dram__bytes_write.sum')
            lines.append(f'{kn:60s} {d["Grid Size"][0].split(",")[0].strip("( "):>6s} {us:9.1f} '
                         f'{val(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"):8.1f} '
                         f'{val(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):6.1f} '
                         f'{(rd + wr) / us / 1e3:9.0f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} '
                         f'{val(d, "lts__t_sector_hit_rate.pct"):7.1f} {val(d, "launch__registers_per_thread"):5.0f} '
                         f'{val(d, "launch__shared_mem_per_block_dynamic") / 1e3:7.1f}')
        open(os.path.join(PROF, f'{TAG}_{name}.txt'), 'w').write('\n'.join(lines) + '\n')
    for src, dst in (('plan_profile_fp16x3.txt', f'{TAG}_plan_profile_fp16x3.txt'), ('plan_profile_c2.txt', f'{TAG}_plan_profile_c2.txt'),
                     ('bench_c2l.log', f'{TAG}_bench_c2l.log'), ('profile_post.txt', f'{TAG}_profile_post.txt'),
                     ('bench_reference.log', f'{TAG}_bench_reference.log'), ('wsi_16384_n1.log', f'{TAG}_wsi_16384_n1.log'),
                     ('bench_nocoal.log', f'{TAG}_bench_direct_epilogue.log'), ('smoke.log', f'{TAG}_smoke.log'),
                     ('parity_report_hifirst.json', f'{TAG}_parity_report_fp16x3_hi_first.json'),
                     ('plan_profile.txt', f'{TAG}_plan_profile.txt'), ('parity_report.json', f'{TAG}_parity_report.json'),
                     ('bench_fp16.log', f'{TAG}_bench_fp16.log'), ('bench_decode_minb3.log', f'{TAG}_bench_decode.log'),
                     ('bench_decode.log', f'{TAG}_bench_decode.log'), ('pytest.log', f'{TAG}_pytest_gpu.log')):
        p = os.path.join(OUT, src)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(PROF, dst))
    print(os.listdir(PROF))


if __name__ == '__main__':
    main()
