#!/bin/bash
# ncu launch list of ONE whole timed window of the bench command (every kernel launch of the 240 h window):
# duration and DRAM bytes per launch.  spin-up 240 h: windows 0-9, warm-up 10-12, timed window = 13.
B="python bench.py --mode ${MODE:-3} --no-e2e --no-cpu-baseline --spinup-hours 240"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
export LGARB_NCU_WINDOW=13 LGARB_NCU_ROUND=-2
$B > gpurun_out/plain_window.log 2>&1 &&
timeout 1500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/window_launches.csv $B > gpurun_out/w.log 2>&1
