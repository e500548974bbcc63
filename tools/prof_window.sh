#!/bin/bash
# ncu launch list of ONE whole timed window of the bench command (every kernel launch of a 240 h window, --steps 5 x 48 h):
# duration and DRAM bytes per launch.  spin-up 240 h in 48 h windows: 0-4, warm-up 5-7, timed window = 8.
B="python bench.py --mode ${MODE:-3} --no-e2e --no-cpu-baseline --no-extras --spinup-hours 240 --steps 5"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
export LGARB_NCU_WINDOW=8 LGARB_NCU_ROUND=-2
$B > gpurun_out/plain_window.log 2>&1 &&
timeout 1500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/window_launches.csv $B > gpurun_out/w.log 2>&1
