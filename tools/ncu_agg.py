"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/ncu_agg.py launches.csv [--seq N]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi, ui, gi, bi = (h.index(x) for x in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Grid Size', 'Block Size'))
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] in ('ns', 'nsecond') else v * 1e3 if r[ui] in ('ms', 'msecond') else v
    name = r[ki].split('(')[0]
    agg[name][0] += 1
    agg[name][1] += v
    seq.append((name, r[gi], r[bi], v))
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':58s} {'n':>5s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:58]:58s} {v[0]:5d} {v[1] / 1e3:10.3f} {v[1] / v[0]:10.1f} {v[1] / tot:7.1%}")
print(f"total {tot / 1e3:.3f} ms over {len(seq)} launches")
if '--seq' in sys.argv:
    n = int(sys.argv[sys.argv.index('--seq') + 1])
    for s in seq[:n]:
        print(f"  {s[0][:40]:40s} grid {s[1]:18s} block {s[2]:14s} {s[3]:10.1f} us")
