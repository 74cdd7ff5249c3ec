"""Per-source-line summary of an ncu report captured with --import-source on:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv ; python profiles/ncu_lines.py X.csv [N]
SASS rows inherit the CUDA line of the row that opened their group."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur, hdr, key = None, None, None
agg = defaultdict(lambda: defaultdict(float)); src = {}
KEYS = ('stall_long_sb', 'stall_short_sb', 'stall_mio', 'stall_lg', 'stall_barrier', 'stall_math', 'stall_wait', 'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_branch_resolving')


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; key = None; continue
    if len(r) >= 60 and r[0] == 'Line No':
        hdr = r; continue
    if len(r) >= 60 and hdr:
        d = dict(zip(hdr, r))
        if r[0].strip().isdigit():
            key = (cur, int(r[0])); src[key] = r[1]
        if key is None:
            continue
        if not r[2].strip():
            continue   # the CUDA line's own summary row: its SASS rows follow
        a = agg[key]
        a['samples'] += f(d['# Samples']); a['inst'] += f(d['Instructions Executed'])
        a['shw'] += f(d['L1 Wavefronts Shared']); a['shx'] += f(d['L1 Wavefronts Shared Excessive'])
        a['gtag'] += f(d.get('L1 Tag Requests Global', 0)); a['loc'] += f(d.get('L2 Theoretical Sectors Local', 0))
        for k in KEYS:
            a[k] += f(d[k])
tot = sum(a['samples'] for a in agg.values())
print('total samples', tot, ' instructions', sum(a['inst'] for a in agg.values()))
for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top_n]:
    st = ' '.join(f"{k[6:]}={a[k]:.0f}" for k in KEYS if a[k] > 0.02 * a['samples'] and a[k] >= 20)
    print(f"{fn[:24]:24s} {ln:4d} {100 * a['samples'] / tot:5.1f}% inst={a['inst']:.0f} shw={a['shw']:.0f} shx={a['shx']:.0f} gtag={a['gtag']:.0f} loc={a['loc']:.0f} | {st} | {src[(fn, ln)].strip()[:80]}")
