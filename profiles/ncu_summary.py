"""Summarise an .ncu-rep (gpurun_out/, scratch) into the text files committed under profiles/:
    python profiles/ncu_summary.py gpurun_out/<name>.ncu-rep > profiles/<name>_ncu.txt"""
import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
keys=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','lts__t_bytes.sum','l1tex__t_bytes.sum']
idx={k:hdr.index(k) for k in keys if k in hdr}
for r in rows[2:]:
    print('-----')
    for k,i in idx.items():
        v=r[i]
        if k=='Kernel Name': v=v[:60]
        print(f"  {k.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio','')[:60]:60s} {v} {rows[1][i]}")
