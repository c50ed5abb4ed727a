"""Side-by-side table of selected metrics for every kernel in an ncu report (the format of profiles/*_ncu_*.txt).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-name-filter ...]
"""
import csv, subprocess, sys

METRICS = [
    'gpu__time_duration.sum', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'lts__t_sector_hit_rate.pct',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
]


def main():
    rep = sys.argv[1]
    filt = sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    name_i = hdr.index('Kernel Name')
    body = [r for r in body if not filt or any(f in r[name_i] for f in filt)]
    def col(h):
        return hdr.index(h) if h in hdr else None
    lines = []
    for label in ('Kernel Name', 'Grid Size', 'Block Size'):
        i = col(label)
        lines.append((label, '', [r[i][:34] for r in body]))
    for mtr in METRICS:
        i = col(mtr)
        if i is None:
            continue
        lines.append((mtr, units[i], [r[i] for r in body]))
    for label, unit, vals in lines:
        print('%-86s %-16s %s' % (label, unit, ' | '.join(vals)))


if __name__ == '__main__':
    main()
