#!/bin/bash
# round 2, call E: a few ncu metrics of the persistent-kernel variants on the soup
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_st.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
for v in 0:16 10 13:16 14:16; do
  VARIANTS=$v timeout -s KILL 600 ncu --metrics $M --clock-control none -k regex:traverse_persistent_kernel -c 1 --csv --log-file gpurun_out/ncu_r2e_${v%%:*}.csv python scripts/trav_sweep.py soup > /dev/null 2>&1
  echo "== $v"; grep -v "^==" gpurun_out/ncu_r2e_${v%%:*}.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print(r['Metric Name'], r['Metric Value'])"
done
