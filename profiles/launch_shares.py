"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share.
    python profiles/launch_shares.py <launches.csv> [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('==')) if r]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1 + skip:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3}.get(r[ui], 1.0)
    name = re.sub(r'\(.*', '', r[ki])
    name = re.sub(r'<unnamed>::', '', name)[:110]
    agg[name][0] += 1
    agg[name][1] += v
total = sum(v for _, v in agg.values())
ours = sum(v for k, (_, v) in agg.items() if 'tsde' in k or 'ew_fast_kernel' in k or 'ew_kernel' in k or 'gen_' in k
           or 'levy' in k or 'bmm_ga' in k or 'cells_' in k or 'bridge_kernel' in k or 'outer_kernel' in k)
print(f"{len(rows) - 1 - skip} launches, {total:.1f} us in total; this library's kernels: {ours:.1f} us = {100 * ours / total:.1f} %")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{100 * v / total:6.2f} %  {v:10.1f} us  {n:6d} x {v / n:8.2f} us  {k}")
