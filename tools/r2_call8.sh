#!/bin/bash
# round 2, eighth GPU call: the GPU suite (binned queues, hourly Update through the stage kernels, forcing frontier of the
# host-buffer verb), what the bins bring (variant builds with 1 and 2 bins), the hourly BMI loop and the end-to-end verbs.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c8_pytest.log
tail -8 gpurun_out/c8_pytest.log
B="timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-extras"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/c8_$name.json 2> gpurun_out/c8_$name.err; }
EXTRA=""
run nbin4 X=1
run nbin1 LGARB_LIBRARY=$PWD/build/liblgarb_nbin1.so
run nbin2 LGARB_LIBRARY=$PWD/build/liblgarb_nbin2.so
run nbin4_dbg LGARB_WAVE_DEBUG=1
run nbin1_dbg LGARB_WAVE_DEBUG=1 LGARB_LIBRARY=$PWD/build/liblgarb_nbin1.so
EXTRA="--hours-per-step 146 --spinup-hours 730"
run L_nbin4 X=1
run L_nbin1 LGARB_LIBRARY=$PWD/build/liblgarb_nbin1.so
# end to end: the host-buffer verbs and the hourly BMI loop, old and new
E="timeout 600 python bench.py --no-cpu-baseline --no-extras"
LGARB_HOST_CHUNK_HOURS=0 $E > gpurun_out/c8_e2e_chunk0.json 2> gpurun_out/c8_e2e_chunk0.err
LGARB_HOST_CHUNK_HOURS=24 $E > gpurun_out/c8_e2e_chunk24.json 2> gpurun_out/c8_e2e_chunk24.err
LGARB_HOST_CHUNK_HOURS=24 LGARB_UPDATE_WAVE_MIN_COLS=100000 $E > gpurun_out/c8_e2e_hourlywave.json 2> gpurun_out/c8_e2e_hourlywave.err
LGARB_HOST_CHUNK_HOURS=8 LGARB_UPDATE_WAVE_MIN_COLS=100000 LGARB_UPDATE_WAVE_SLOTS=1000000 $E > gpurun_out/c8_e2e_hourlywave_1m.json 2> gpurun_out/c8_e2e_hourlywave_1m.err
for f in gpurun_out/c8_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); e=d.get('e2e') or {}
    print(' value %.4g launches %s failed %s e2e %s streamed %s hourly %s'%(d['value'], d['gpu_launches'], d['failed_columns'], e.get('value'), (e.get('streamed') or {}).get('value'), (e.get('hourly_bmi_loop') or {}).get('value')))
except Exception as ex: print(' FAILED', ex)
"; done
