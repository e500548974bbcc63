#!/bin/bash
# round 2, eleventh GPU call (the tenth ran a build whose queue counters beyond 256 were never initialised: 24 bins x 12 queues = 288): GPU suite with 6 x 4 bins (correction-type flag), its effect, the hourly loop with the fast search path
# in the finishing kernel, the stress variant of config 3 at 1 M columns.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c11_pytest.log
tail -4 gpurun_out/c11_pytest.log
grep -q "pytest rc 0" gpurun_out/c11_pytest.log || { echo "GPU suite not green: stopping here"; grep -E "^FAILED|Error" gpurun_out/c11_pytest.log | head -20; exit 1; }
B="timeout 150 python bench.py --no-cpu-baseline --no-e2e --no-extras"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/c11_$name.json 2> gpurun_out/c11_$name.err; }
EXTRA=""
run n6f4 X=1
run n6f2 LGARB_LIBRARY=$PWD/build/liblgarb_n6f2.so
run n6f4_dbg LGARB_WAVE_DEBUG=1
EXTRA="--hours-per-step 146 --spinup-hours 730"
run L_n6f4 X=1
run L_n6f2 LGARB_LIBRARY=$PWD/build/liblgarb_n6f2.so
H="timeout 150 python tools/hourly_profile.py"
LGARB_UPDATE_WAVE_MIN_COLS=100000 $H > gpurun_out/c11_hourly_wave_ff1.json 2> gpurun_out/c11_hourly_wave_ff1.err
LGARB_UPDATE_WAVE_MIN_COLS=100000 LGARB_FINISH_FF=0 $H > gpurun_out/c11_hourly_wave_ff0.json 2> gpurun_out/c11_hourly_wave_ff0.err
LGARB_UPDATE_WAVE_MIN_COLS=100000 LGARB_WAVE_FINISH_BELOW=40000 $H > gpurun_out/c11_hourly_wave_ff1_40k.json 2> gpurun_out/c11_hourly_wave_ff1_40k.err
LGARB_UPDATE_WAVE_MIN_COLS=100000 LGARB_WAVE_FF_BELOW=300000 $H > gpurun_out/c11_hourly_wave_ffrounds.json 2> gpurun_out/c11_hourly_wave_ffrounds.err
LGARB_UPDATE_WAVE_MIN_COLS=100000 LGARB_WAVE_DEBUG=1 $H --hours 3 > gpurun_out/c11_hourly_dbg.json 2> gpurun_out/c11_hourly_dbg.err
cat gpurun_out/c11_hourly_wave_*.json
grep "wave\]" gpurun_out/c11_hourly_dbg.err | tail -8
timeout 300 python tools/bench_stress.py --columns 1000000 --hours 12 > gpurun_out/c11_stress_1m.json 2> gpurun_out/c11_stress_1m.err
cat gpurun_out/c11_stress_1m.json
for f in gpurun_out/c11_*n6*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
