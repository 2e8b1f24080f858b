"""Summarise an `ncu --page raw --csv` dump: the handful of metrics DESIGN.md / bench.py cite."""
import csv, sys
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__t_sectors_srcunit_tex_op_read.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.per_cycle_active','smsp__warps_eligible.avg.per_cycle_active',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg.per_second','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','smsp__inst_executed_op_shfl.sum' if False else 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:100], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:88s} {r[i]:>18s} {units[i]}")
