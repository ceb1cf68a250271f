#!/usr/bin/env python
"""Refresh the TABLES of profiles/r2_ncu.md from the evidence files of a final run (the prose between them is edited by hand):

    ncu -i gpurun_out/prof_TAG.ncu-rep --page raw --csv > /tmp/raw.csv
    ncu -i gpurun_out/prof_TAG.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:prb_setup > /tmp/src_setup.csv
    python tools/make_profile_md.py /tmp/raw.csv /tmp/src_setup.csv

Reads profiles/r2_bench*.json and profiles/r2_launch_summary.txt; rewrites the sections between the '## ' headings
'Bench lines', 'Launch list', 'Full-set metrics' and 'Setup kernel by phase' up to their closing marker lines."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, 'profiles')


def J(f):
    return json.load(open(os.path.join(P, f)))


def fmt(v):
    return format(round(v), ',').replace(',', ' ')


def row(name, d):
    r = d['roofline']
    reset = ('%.0f ms / %d rounds' % (d['reset']['full_batch_ms'], d['reset']['rounds'])) if d.get('reset') else ''
    return '| %s | %s | %s | %.1f ms (setup %.1f, solve %.1f) | %s |' % (name, fmt(d['value']), fmt(d['e2e']['value']), d['ms_per_step'],
                                                                          r['setup_kernels_ms'], r['pgs_kernels_ms'], reset)


def between(s, start, end, new):
    a = s.index(start)
    b = s.index(end, a)
    return s[:a] + new + s[b:]


def main():
    raw_csv, src_csv = sys.argv[1], sys.argv[2]
    b, jmp, npre, b8 = J('r2_bench.json'), J('r2_bench_jump.json'), J('r2_bench_nopreroll.json'), J('r2_bench_8192.json')
    pk, rc, ref = J('r2_bench_pick16384.json'), J('r2_bench_reach4096.json'), J('r2_bench_ref.json')
    w8, s8, s4, s2 = J('r2_bench_8gpu_weak.json'), J('r2_bench_8gpu_strong.json'), J('r2_bench_4gpu_strong.json'), J('r2_bench_2gpu_strong.json')
    bench = '\n'.join([
        '| workload | env-steps/s (device) | end to end | step pipeline per env step | full-batch reset |', '|---|---|---|---|---|',
        row('UR5PlayAbsRPY1Obj-v0, 65 536 envs, scripted steady state incl. drawer / door handles (headline)', b),
        row('same, + 5 % Random-stream jump tail (`--jump-frac 0.05`)', jmp),
        row('same, no pre-roll (`--preroll 0`: states right after reset)', npre),
        row('UR5PlayAbsRPY1Obj-v0, 8 192 envs (BASELINE config 4)', b8),
        row('pandaPick-v0, 16 384 envs (config 3)', pk),
        row('UR5Reach-v0, 4 096 envs (config 2)', rc),
        '| CPU oracle port, 16 host cores (`--impl reference`, same scripted stream and pre-roll) | %s | | | |' % fmt(ref['value']), '',
        'The end-to-end leg replays the device-timed steps from the same saved states through `VecPlayEnv.step` (host actions in, a fresh host',
        'array of the whole observation block out, every step).  `capacity_overflow_env_steps` is 0 in every line except the jump workload',
        '(1 env-step in 1.3 M).  8 GPUs, one process per GPU, env-index sharding, no data-path collective: weak scaling (65 536 envs per GPU)',
        '**%.2f M env-steps/s** (%.2f x one GPU; end to end %.2f M); strong scaling of a fixed 65 536-env batch over 2 / 4 / 8 GPUs:' % (
            w8['value'] / 1e6, w8['value'] / b['value'], w8['e2e']['value'] / 1e6),
        '%.2f / %.2f / %.2f M env-steps/s (%.1f / %.1f / %.1f ms per step: the latency floor of a substep, the 50-sweep chain of' % (
            s2['value'] / 1e6, s4['value'] / 1e6, s8['value'] / 1e6, s2['ms_per_step'], s4['ms_per_step'], s8['ms_per_step']),
        'the largest arm islands, does not shrink with the batch).', '', ''])
    table = subprocess.check_output([sys.executable, os.path.join(ROOT, 'tools', 'ncu_summary.py'), raw_csv]).decode()
    lines = table.split('\n')
    lines[0] = ('| metric | `prb_setup_kernel<12>` | `prb_pgs_arm_kernel` class 4 (1 env/warp) | class 3 (8 envs/warp) | class 2 (8) | class 1 (8) | '
                'class 0 (32) | `prb_pgs_joint_kernel<12>` | `prb_pgs_free_kernel` |')
    table = '\n'.join(lines)
    phase = subprocess.check_output([sys.executable, os.path.join(ROOT, 'tools', 'ncu_phase_breakdown.py'), src_csv]).decode()
    phase = '\n'.join(phase.split('\n')[:22])
    launch = open(os.path.join(P, 'r2_launch_summary.txt')).read().strip()
    path = os.path.join(P, 'r2_ncu.md')
    s = open(path).read()
    s = between(s, '| workload | env-steps/s (device)', 'Round 1 (v27) on its own', bench)
    s = between(s, '## Launch list', 'In the real step the solver kernels', '## Launch list: share of one env step (serialised by ncu)\n\n```\n%s\n```\n\n' % launch)
    s = between(s, '## Full-set metrics', '## Setup kernel by phase', '## Full-set metrics of one substep (ncu --set full)\n\n%s\n' % table)
    s = between(s, '## Setup kernel by phase', '## Reading',
                '## Setup kernel by phase (source-level attribution of the same capture, `tools/ncu_phase_breakdown.py`)\n\n```\n%s\n```\n\n' % phase)
    open(path, 'w').write(s)
    print('updated', path)


if __name__ == '__main__':
    main()
