#!/usr/bin/env python
"""Phase breakdown of a kernel from an ncu source-page export.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-id ::regex:NAME:1 > src.csv
    python tools/ncu_phase_breakdown.py src.csv

Every SASS instruction row carries executed-instruction and stall-sample counts.  Rows are attributed to the device
function (phase_*, box_box, side_row, ...) whose source lines produced them; instructions that come from the small
inlined helpers of prb_device.h are attributed to the function of the nearest preceding instruction (by address) that
maps to prb_kernels.cuh / prb_stream.cuh / prb_reset.cuh.
"""
import csv
import os
import re
import sys
from collections import defaultdict

CSRC = os.environ.get('PRB_CSRC') or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'roboticsplayroompybullet_b200', 'csrc')
FILES = ['prb_kernels.cuh', 'prb_stream.cuh', 'prb_reset.cuh']


def function_spans(path):
    spans = []
    for i, l in enumerate(open(path).read().split('\n'), 1):
        if l.startswith(('PRB_D ', 'PRB_DN ', '__global__', 'static PRB_D')) or re.match(r'^\s*(static )?PRB_D ', l):
            m = re.search(r'\b([A-Za-z_][A-Za-z0-9_]*)\s*\(', l)
            if l.startswith('__global__'):
                m = re.search(r'\b(prb_[a-z_]+kernel)\b', l) or m
            if m:
                spans.append((i, m.group(1)))
    return spans


def main():
    spans = {f: function_spans(os.path.join(CSRC, f)) for f in FILES}

    def func_of(f, line):
        name = '?'
        for first, n in spans[f]:
            if first <= line:
                name = n
            else:
                break
        return name

    rows, hdr, fpath, line = [], None, None, None
    for row in csv.reader(open(sys.argv[1], newline='')):
        if not row:
            continue
        if row[0] == 'File Path':
            fpath = os.path.basename(row[1])
            continue
        if row[0] in ('Function Name', 'Kernel Name'):
            continue
        if row[0] == 'Line No':
            hdr = row
            continue
        if hdr is None:
            continue
        if row[0] != '':
            try:
                line = int(row[0])
            except ValueError:
                line = None
            continue
        ia, ii, isamp, it = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
        if len(row) <= ii or not row[ia].startswith('0x'):
            continue
        try:
            rows.append((int(row[ia], 16), fpath, line, int(row[ii]), int(row[isamp]), int(row[it])))
        except ValueError:
            pass
    rows.sort()
    seen, uniq = set(), []
    for r in rows:
        if r[0] not in seen:
            seen.add(r[0])
            uniq.append(r)
    inst, samp, thr, static = defaultdict(int), defaultdict(int), defaultdict(int), defaultdict(int)
    cur = 'prologue'
    for addr, f, ln, ni, ns, nt in uniq:
        if f in spans and ln:
            cur = func_of(f, ln)
        inst[cur] += ni
        static[cur] += 1
        samp[cur] += ns
        thr[cur] += nt
    ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
    print('%d SASS instructions, %.3e warp instructions executed, %d samples' % (len(uniq), ti, ts))
    for p in sorted(inst, key=lambda p: -samp[p]):
        print('   %-26s inst %6.2f%%   samples %6.2f%%   active lanes %4.1f   SASS %5d' % (p, 100.0 * inst[p] / ti, 100.0 * samp[p] / ts, thr[p] / max(inst[p], 1), static[p]))


if __name__ == '__main__':
    main()
