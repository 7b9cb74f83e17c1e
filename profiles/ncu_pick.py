"""Prints a selection of metrics from `ncu -i X.ncu-rep --page raw --csv` output (one column per launch)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__sass_average_branch_targets_threads_uniform.pct',
        'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2:]
print("kernels:", [v[hdr.index('Kernel Name')][:60] for v in vals])
for w in WANT + sys.argv[2:]:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:86s} {units[i]:14s} {[v[i] for v in vals]}")
