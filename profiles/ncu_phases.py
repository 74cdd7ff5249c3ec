"""Samples of an ncu source-page export (--print-source cuda,sass) in SASS ADDRESS order, cut into phases at the
barrier instructions (BAR.SYNC, UCGABAR_WAIT, SYNCS...TRYWAIT loops): python profiles/ncu_phases.py X_source.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
sass = {}
line_of = {}
cur_line = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) >= 60 and r[0] == 'Line No':
        hdr = r
        continue
    if len(r) >= 60 and r[0] == 'Address':   # SASS-only export
        hdr = r
        continue
    if hdr and hdr[0] == 'Address' and len(r) >= 60 and r[0].startswith('0x'):
        d = dict(zip(hdr, r))
        sass[int(r[0], 16)] = (r[1].strip(), float(d['# Samples']), float(d['Instructions Executed']), None)
        continue
    if hdr and len(r) >= 60:
        if r[0].strip().isdigit():
            cur_line = (cur_file, int(r[0]))
        if r[2].startswith('0x'):
            d = dict(zip(hdr, r))
            try:
                sass[int(r[2], 16)] = (r[3].strip(), float(d['# Samples']), float(d['Instructions Executed']), cur_line)
            except ValueError:
                pass
addrs = sorted(sass)
tot = sum(s[1] for s in sass.values())
acc_s = acc_i = 0.0
start = addrs[0]
lines = set()
print(f"total samples {tot:.0f}")
for a in addrs:
    op, smp, inst, ln = sass[a]
    acc_s += smp
    acc_i += inst
    if ln and 'cluster' in ln[0] or ln and 'xline_kernel' in ln[0]:
        lines.add(ln[1])
    if any(t in op for t in ('BAR.SYNC', 'UCGABAR_WAIT', 'TRYWAIT', 'EXIT')):
        if acc_s > 0.003 * tot:
            print(f"{a - addrs[0]:6x} {op[:40]:40s} {100 * acc_s / tot:5.1f}%  inst {acc_i / 1e6:8.1f} M  lines {min(lines) if lines else 0}-{max(lines) if lines else 0}")
        acc_s = acc_i = 0.0
        lines = set()
