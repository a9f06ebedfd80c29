"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share for the
LAST `--last N` launches (one bench step)."""
import csv, sys, collections, re
path = sys.argv[1]
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        grid = r.get('Grid Size', '')
        rows.append((name, ns, grid))
if last:
    rows = rows[-last:]
agg = collections.OrderedDict()
for name, ns, grid in rows:
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + ns)
tot = sum(t for _, t in agg.values())
print(f'launches {len(rows)} total {tot/1e6:.3f} ms')
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{t/1e6:9.3f} ms {100*t/tot:5.1f}%  n={c:5d}  avg {t/c/1e3:9.1f} us  {name}')
