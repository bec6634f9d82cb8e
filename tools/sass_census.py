"""Blackwell-native instruction census of the built library: `cuobjdump -sass` per kernel, counting the mnemonics that prove the
tensor path is tcgen05 / TMA (UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = kind::f8f6f4, UTMALDG = TMA tensor load, LDTM =
tcgen05.ld, UTCBAR = tcgen05.commit, LDGSTS = cp.async) and that no mma.sync (HMMA) is left.  CPU only.
Usage: python tools/sass_census.py > profiles/rNN_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'celldetection_b200', 'libcpn_b200.so')
WANT = ('UTCHMMA', 'UTCQMMA', 'UTCOMMA', 'UTCIMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'LDGSTS', 'HMMA', 'IMMA', 'ELECT',
        'MATCH', 'REDUX', 'ATOMS', 'ATOMG', 'RED')


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    per, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r'\(.*', '', name).replace('void ', '').replace('cpn::', '')
            per[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.search(r'/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1).split('.')[0]
            if op in WANT:
                per[name][op] += 1
    print('# cuobjdump -sass celldetection_b200/libcpn_b200.so (sm_100a): Blackwell-native instruction counts per kernel')
    print('# UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = tcgen05.mma kind::f8f6f4 (the e4m3 correction pass), UTMALDG = TMA tensor load,')
    print('# LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit, LDGSTS = cp.async, MATCH / REDUX = __match_any_sync /')
    print('# __reduce_*_sync (histogram-free merging in the label statistics), ATOMS = shared-memory atomics (histograms, radix sort).')
    total = collections.Counter()
    for k, c in per.items():
        if c:
            print(f'{k[:58]:60s}' + '  '.join(f'{op}={n}' for op, n in sorted(c.items())))
            total.update(c)
    print('# total: ' + '  '.join(f'{op}={n}' for op, n in sorted(total.items())) + f'   kernels in the library: {len(per)}')
    assert total['HMMA'] == 0 and total['IMMA'] == 0, 'mma.sync found'


if __name__ == '__main__':
    main()
