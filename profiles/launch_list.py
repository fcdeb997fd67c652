"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <bench command>`):
kernel share of the captured launches.  Per-launch times under ncu are cold-cache and serialised: the SHARE is what is compared with
bench.py's live CUDA-event breakdown, not the absolute.   python profiles/launch_list.py X.csv > summary.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki])
    name = re.sub(r'^void ', '', name)
    v = float(r[vi].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1.0)
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + v)
total = sum(t for _, t in tot.values())
ours = sum(t for k, (_, t) in tot.items() if k.startswith('dsb::'))
print(f'# {sys.argv[1]}: {sum(n for n, _ in tot.values())} launches, {total / 1e3:.2f} ms under ncu; kernels of this repo (dsb::*): {100 * ours / total:.1f} % of the time')
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{k[:90]:90s} n={n:5d}  {t / 1e3:10.3f} ms  {100 * t / total:5.1f} %')
