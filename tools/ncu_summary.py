#!/usr/bin/env python
"""Summarise an ncu report for profiles/: per-kernel headline metrics (raw page) as markdown.

    ncu -i gpurun_out/prof_X.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv > profiles/<round>_<tag>_ncu.md
"""
import csv
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
    ('launch__occupancy_limit_shared_mem', 'occupancy limit (smem), blocks'),
    ('launch__occupancy_limit_registers', 'occupancy limit (regs), blocks'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('smsp__warps_active.avg.per_cycle_active', 'warps active / scheduler'),
    ('smsp__warps_eligible.avg.per_cycle_active', 'warps eligible / scheduler'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active threads / instruction'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe %'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('lts__t_bytes.sum', 'L2 bytes'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'shared wavefronts'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared bank conflicts'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait / issue'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier / issue'),
    ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'stall branch_resolving / issue'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall not_selected / issue'),
    ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'stall no_instruction / issue'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math_pipe_throttle / issue'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall lg_throttle / issue'),
    ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall mio_throttle / issue'),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = rows[2:]
    print('| metric | ' + ' | '.join('`%s`' % r[idx['Kernel Name']].split('(')[0].replace('void ', '')[:48] for r in kernels) + ' |')
    print('|---|' + '---|' * len(kernels))
    for key, label in WANT:
        if key not in idx:
            continue
        u = units[idx[key]]
        vals = []
        for r in kernels:
            v = r[idx[key]]
            try:
                f = float(v.replace(',', ''))
                v = ('%.4g' % f) if abs(f) < 1e6 else ('%.4e' % f)
            except ValueError:
                pass
            vals.append(v + (' ' + u if u else ''))
        print('| %s | %s |' % (label, ' | '.join(vals)))


if __name__ == '__main__':
    main()
