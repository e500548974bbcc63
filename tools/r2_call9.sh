#!/bin/bash
# round 2, ninth GPU call: finer bins (variant builds), search quantum / pairs with bins, where an hour of the BMI loop goes.
mkdir -p gpurun_out
B="timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-extras"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/c9_$name.json 2> gpurun_out/c9_$name.err; }
EXTRA=""
run n4f1 X=1
run n4f2 LGARB_LIBRARY=$PWD/build/liblgarb_n4f2.so
run n6f1 LGARB_LIBRARY=$PWD/build/liblgarb_n6f1.so
run n6f2 LGARB_LIBRARY=$PWD/build/liblgarb_n6f2.so
run n4f1_q96 LGARB_WAVE_QUANTUM=96
run n4f1_q128_p3 LGARB_WAVE_QUANTUM=128 LGARB_WAVE_PAIRS=3
run n4f1_p6 LGARB_WAVE_PAIRS=6
run n4f1_dbg LGARB_WAVE_DEBUG=1
EXTRA="--hours-per-step 146 --spinup-hours 730"
run L_n4f1 X=1
run L_n4f2 LGARB_LIBRARY=$PWD/build/liblgarb_n4f2.so
run L_n6f2 LGARB_LIBRARY=$PWD/build/liblgarb_n6f2.so
timeout 300 python tools/hourly_profile.py > gpurun_out/c9_hourly_pair.json 2> gpurun_out/c9_hourly_pair.err
LGARB_UPDATE_WAVE_MIN_COLS=100000 timeout 300 python tools/hourly_profile.py > gpurun_out/c9_hourly_wave.json 2> gpurun_out/c9_hourly_wave.err
LGARB_UPDATE_WAVE_MIN_COLS=100000 LGARB_WAVE_DEBUG=1 timeout 300 python tools/hourly_profile.py --hours 4 > gpurun_out/c9_hourly_wave_dbg.json 2> gpurun_out/c9_hourly_wave_dbg.err
cat gpurun_out/c9_hourly_pair.json gpurun_out/c9_hourly_wave.json
grep "wave\]" gpurun_out/c9_hourly_wave_dbg.err | tail -12
for f in gpurun_out/c9_*n*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
