#!/bin/bash
# ncu capture of one round of the wavefront scheduler inside the bench's timed window.
# usage: prof_round.sh ROUND [full]    (spin-up 240 h in 48 h windows: 0-4, warm-up 5-7, timed window = 8)
R=${1:-5}
B="python bench.py --mode ${MODE:-3} --no-e2e --no-cpu-baseline --no-extras --spinup-hours 240 --steps 5"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
export LGARB_NCU_WINDOW=8 LGARB_NCU_ROUND=$R
$B > gpurun_out/plain_round.log 2>&1 || exit 1
if [ "$2" = full ]; then
  timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name regex:lgar_wave_ -o gpurun_out/round${R}_full $B > gpurun_out/rf.log 2>&1
else
  timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/round${R}_metrics.csv $B > gpurun_out/r.log 2>&1
fi
