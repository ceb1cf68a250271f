#!/usr/bin/env python
"""Phase breakdown of the fused step kernel from an ncu source-page export.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_phase_breakdown.py src.csv [kernel-name-substring]

Every SASS instruction row carries executed-instruction and stall-sample counts.  Rows are attributed
to the device function (phase_*, box_box, ...) whose source lines produced them; instructions that
come from the small inlined helpers of prb_device.h are attributed to the phase of the nearest
preceding instruction (by address) that maps to prb_kernels.cuh.
"""
import csv
import re
import sys
from collections import defaultdict

KERNELS = 'roboticsplayroompybullet_b200/csrc/prb_kernels.cuh'


def function_spans(path):
    """[(first_line, name)] of the top-level device functions of prb_kernels.cuh."""
    spans = []
    pat = re.compile(r'^(?:PRB_DN?|__global__|static|template)?.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(')
    lines = open(path).read().split('\n')
    for i, l in enumerate(lines, 1):
        if l.startswith(('PRB_D ', 'PRB_DN ', '__global__')):
            m = re.search(r'\b([A-Za-z_][A-Za-z0-9_]*)\s*\(', l.split(')', 1)[0] if l.startswith('__global__') else l)
            if l.startswith('__global__'):
                m = re.search(r'\b(prb_[a-z_]+kernel)\b', l) or m
            if m:
                spans.append((i, m.group(1)))
    return spans


def main():
    src = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ''
    spans = function_spans(KERNELS)

    def func_of(line):
        name = '?'
        for first, n in spans:
            if first <= line:
                name = n
            else:
                break
        return name

    sections, cur = [], None
    fpath = None
    for row in csv.reader(open(src, newline='')):
        if not row:
            continue
        if row[0] == 'File Path':
            fpath = row[1]
            continue
        if row[0] == 'Function Name':
            cur = {'kernel': row[1], 'file': fpath, 'rows': [], 'hdr': None}
            sections.append(cur)
            continue
        if row[0] == 'Line No':
            cur['hdr'] = row
            continue
        if cur is not None and cur['hdr'] is not None:
            cur['rows'].append(row)
    # instruction rows: (addr, file, line, inst, samples)
    by_kernel = defaultdict(list)
    for s in sections:
        h = s['hdr']
        ia, ii, isamp = h.index('Address'), h.index('Instructions Executed'), h.index('# Samples')
        line = None
        for r in s['rows']:
            if r[0] != '':
                try:
                    line = int(r[0])
                except ValueError:
                    line = None
                continue
            if len(r) <= ii or not r[ia].startswith('0x'):
                continue
            try:
                by_kernel[s['kernel']].append((int(r[ia], 16), s['file'], line, int(r[ii]), int(r[isamp])))
            except ValueError:
                pass
    for k, rows in by_kernel.items():
        if want and want not in k:
            continue
        rows.sort()
        seen, uniq = set(), []
        for r in rows:
            if r[0] in seen:
                continue
            seen.add(r[0])
            uniq.append(r)
        phase_inst, phase_samp = defaultdict(int), defaultdict(int)
        cur_phase = 'prologue'
        for addr, f, line, inst, samp in uniq:
            if f and f.endswith('prb_kernels.cuh') and line:
                cur_phase = func_of(line)
            phase_inst[cur_phase] += inst
            phase_samp[cur_phase] += samp
        ti, ts = sum(phase_inst.values()) or 1, sum(phase_samp.values()) or 1
        print('==', k[:110])
        print('   %d SASS instructions, %.3e warp-instructions executed, %d samples' % (len(uniq), ti, ts))
        for p in sorted(phase_inst, key=lambda p: -phase_samp[p]):
            print('   %-22s inst %6.2f%%   samples %6.2f%%' % (p, 100.0 * phase_inst[p] / ti, 100.0 * phase_samp[p] / ts))


if __name__ == '__main__':
    main()
