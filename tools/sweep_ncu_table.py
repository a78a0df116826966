"""Prints the markdown table of profiles/r02_ncu_summary.md section 3 from profiles/r02_sweep_kernel.ncu-rep (ncu -i ... --page raw --csv)."""
import csv, subprocess, sys, json
rep='/root/repo/profiles/r02_sweep_kernel.ncu-rep'
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(out.splitlines())); h=r[0]; v=dict(zip(h,r[2])); u=dict(zip(h,r[1]))
keys=['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__shared_mem_per_block_dynamic',
'dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed',
'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum',
'smsp__average_warp_latency_per_inst_issued.ratio','sass__inst_executed_local_loads',
'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
print("| metric | value | unit |\n|---|---|---|")
for k in keys:
    if k in v: print(f"| `{k}` | {v[k]} | {u.get(k,'')} |")
