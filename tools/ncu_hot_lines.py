"""The lines of a kernel that collect the most warp-stall samples, from `ncu -i report --page source --csv` output:
    ncu -i gpurun_out/x.ncu-rep --page source --csv --launch-skip K --launch-count 1 > /tmp/src.csv; python tools/ncu_hot_lines.py /tmp/src.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
data = []
for r in rows:
    if '# Samples' in r:
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    data.append(r)
iS, iSrc, iI = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
num = lambda v: int(v) if v.strip().isdigit() else 0
tot = sum(num(r[iS]) for r in data)
print('total samples', tot, 'rows', len(data))
for r in sorted(data, key=lambda r: -num(r[iS]))[:top_n]:
    why = sorted(((num(r[i]), hdr[i][6:]) for i in stalls), reverse=True)[:2]
    print('%6d %5.1f%% inst=%-9s %-28s %s' % (num(r[iS]), 100.0 * num(r[iS]) / max(tot, 1), r[iI],
                                             ' '.join('%s=%d' % (n, c) for c, n in why if c), r[iSrc][:110]))
