#!/usr/bin/env python
"""Hottest source lines of one kernel from an ncu source-page CSV export.
    python tools/ncu_lines.py src.csv KERNEL_SUBSTR [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], newline='')))
want, topn = sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30
cur = hdr = fpath = None
out = {}
stall_cols = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fpath = r[1]; continue
    if r[0] == 'Function Name': cur = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if cur and want in cur and r[0] != '' and hdr:
        try: line = int(r[0])
        except ValueError: continue
        key = (fpath.split('/')[-1], line)
        if key in out: continue
        try: samp = int(r[hdr.index('# Samples')]); inst = int(r[hdr.index('Instructions Executed')])
        except ValueError: continue
        st = {}
        for name in ('stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_lg', 'stall_mio', 'stall_branch_resolving', 'stall_no_inst', 'stall_barrier', 'stall_math', 'stall_dispatch', 'stall_membar', 'stall_drain'):
            if name in hdr:
                try: st[name] = int(r[hdr.index(name)])
                except ValueError: pass
        out[key] = (samp, inst, r[1][:90], st)
tot = sum(v[0] for v in out.values()) or 1; ti = sum(v[1] for v in out.values()) or 1
print('total samples', tot, 'warp instructions', ti)
for k, v in sorted(out.items(), key=lambda x: -x[1][0])[:topn]:
    top = sorted(v[3].items(), key=lambda x: -x[1])[:2]
    print('%-16s %4d samp %5.1f%% inst %5.1f%% %-38s %s' % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / ti, ' '.join('%s=%d' % (a.replace('stall_', ''), b) for a, b in top if b), v[2]))
