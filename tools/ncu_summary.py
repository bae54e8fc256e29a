"""Summaries of .ncu-rep captures for profiles/: `python tools/ncu_summary.py raw <rep> <out>` (selected raw-page metrics of every
kernel in the report, one block per kernel), `... table <rep> <out>` (one line per kernel), `... stalls <rep> <out>` (per-source-line
warp-stall samples through tools/ncu_lines.py)."""
import csv, io, subprocess, sys

KEEP = ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum', 'smsp__inst_executed.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio')


def raw_rows(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


mode, rep, dst = sys.argv[1:4]
if mode in ('raw', 'table'):
    hdr, units, rows = raw_rows(rep)
    ki = hdr.index('Kernel Name')
    with open(dst, 'w') as f:
        if mode == 'raw':
            for r in rows:
                f.write('%-110s %s\n' % ('Kernel Name', r[ki]))
                for name in sorted(KEEP):
                    if name in hdr:
                        j = hdr.index(name)
                        f.write('%-110s %-12s %s\n' % (name, units[j], r[j]))
                f.write('\n')
        else:
            cols = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
                    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread']
            cols = [c for c in cols if c in hdr]
            f.write(' | '.join(['Kernel Name'] + cols) + '\n')
            f.write(' | '.join([''] + [units[hdr.index(c)] for c in cols]) + '\n')
            for r in rows:
                f.write(' | '.join([r[ki][:60]] + [r[hdr.index(c)] for c in cols]) + '\n')
elif mode == 'stalls':
    tmp = dst + '.src.csv'
    with open(tmp, 'w') as f:
        subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], stdout=f, check=True)
    with open(dst, 'w') as f:
        subprocess.run([sys.executable, 'tools/ncu_lines.py', tmp, '0.005'], stdout=f, check=True)
    import os; os.remove(tmp)
