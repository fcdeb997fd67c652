"""SASS evidence for the tcgen05 / TMEM / TMA claims of DESIGN.md section 4: per kernel of the shipped library, how many of the
Blackwell tensor-core / tensor-memory / bulk-tensor-copy instructions cuobjdump finds.  Runs without a GPU.

    python profiles/sass_summary.py > profiles/r02/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'diff-sampler_b200', 'libdiffsampler_b200.so')
MNEMONICS = ['UTCHMMA', 'UTCQMMA', 'UTCHMMA.2CTA', 'UTCQMMA.2CTA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'SYNCS', 'HMMA', 'IMMA', 'MUFU.EX2',
             'LDG.E.128', 'STG.E.128', 'LDS', 'STS']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    cur = None
    counts = collections.OrderedDict()
    for ln in out.splitlines():
        m = re.search(r'Function : (\S+)', ln)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
        if not m:
            continue
        op = m.group(1)
        counts[cur]['_total'] += 1
        for k in MNEMONICS:
            base = k.split('.')[0]
            if k in ('UTCHMMA', 'UTCQMMA', 'UTCHMMA.2CTA', 'UTCQMMA.2CTA'):
                if op.split('.')[0] == base and (('.2CTA' in op) == k.endswith('.2CTA')):
                    counts[cur][k] += 1
            elif op == k or op.startswith(k + '.'):
                counts[cur][k] += 1
    print(f'# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a); instruction counts per kernel')
    cols = ['UTCHMMA', 'UTCQMMA', 'UTCHMMA.2CTA', 'UTCQMMA.2CTA', 'UTCBAR', 'LDTM', 'UTMALDG', 'SYNCS', 'HMMA', 'MUFU.EX2', 'LDG.E.128', 'STG.E.128']
    print(f'{"kernel":72s} {"instrs":>7s} ' + ' '.join(f'{c:>12s}' for c in cols))
    for k, c in counts.items():
        name = k if len(k) <= 72 else k[:69] + '...'
        print(f'{name:72s} {c["_total"]:7d} ' + ' '.join(f'{c[x]:12d}' for x in cols))
    tc = [k for k, c in counts.items() if c['UTCHMMA'] or c['UTCQMMA']]
    print(f'\n# kernels issuing tcgen05.mma (UTCHMMA = kind::f16, UTCQMMA = kind::f8f6f4): {len(tc)}: ' + ', '.join(tc))
    print('# legacy warp-level tensor instructions (HMMA / IMMA) in the library: ' + str(sum(c['HMMA'] + c['IMMA'] for c in counts.values())))


if __name__ == '__main__':
    sys.exit(main())
