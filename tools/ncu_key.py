"""Print the key metrics of every kernel in an .ncu-rep (ncu --set full) capture.
    python tools/ncu_key.py gpurun_out/full_gather.ncu-rep [extra metric substrings]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name','Grid Size','Block Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum',
 'dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum','sm__inst_executed_pipe_tensor.sum',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg','smsp__cycles_active.avg','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
extra = sys.argv[2:]  # substrings
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('='*100)
    for k in keys:
        if k in d: print('%-90s %-12s %s' % (k, units[hdr.index(k)], d[k]))
    for s in extra:
        for h in hdr:
            if s in h and h not in keys: print('%-90s %-12s %s' % (h, units[hdr.index(h)], d[h]))
