"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line:
   python scripts/ncu_lines.py dump.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, data = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = {h: j for j, h in enumerate(r)}
        ie, isamp = hdr.get('Instructions Executed'), hdr.get('# Samples', hdr.get('Warp Stall Sampling (All Samples)'))
    elif hdr and r[0].isdigit():
        def num(j):
            try:
                return int(r[j])
            except (ValueError, IndexError, TypeError):
                return 0
        data.append((cur_file, int(r[0]), r[1], num(ie) if ie is not None else 0, num(isamp)))
ti, ts = sum(d[3] for d in data) or 1, sum(d[4] for d in data) or 1
print(f'total warp instructions {ti}, samples {ts}')
for d in sorted(data, key=lambda d: -d[4])[:top]:
    print(f'{d[0]:16s}:{d[1]:4d} inst {d[3]:9d} {d[3] / ti:6.3f}  samples {d[4]:6d} {d[4] / ts:6.3f} | {d[2][:110]}')
