#!/bin/bash
# round 2, seventh GPU call: the GPU suite (frozen-snapshot function tests, finishing-kernel variants), the records-per-warp and
# hand-over sweep of the shared-memory finishing kernel, the default bench line with every extra, BASELINE config 3 at size.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c7_pytest.log
tail -8 gpurun_out/c7_pytest.log
B="timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-extras"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/c7_$name.json 2> gpurun_out/c7_$name.err; }
EXTRA=""
run spw16 LGARB_FINISH_SPW=16
run spw8 LGARB_FINISH_SPW=8
run spw4 LGARB_FINISH_SPW=4
run spw2 LGARB_FINISH_SPW=2
run spw8_32k LGARB_FINISH_SPW=8 LGARB_WAVE_FINISH_BELOW=32768
run spw4_32k LGARB_FINISH_SPW=4 LGARB_WAVE_FINISH_BELOW=32768
run spw2_9k LGARB_FINISH_SPW=2 LGARB_WAVE_FINISH_BELOW=9472
run spw4_q256 LGARB_FINISH_SPW=4 LGARB_FINISH_QUANTUM=256
run spw4_ff LGARB_FINISH_SPW=4 LGARB_FINISH_FF=1
EXTRA="--hours-per-step 146 --spinup-hours 730"
run L_spw16 LGARB_FINISH_SPW=16
run L_spw4 LGARB_FINISH_SPW=4
run L_spw2 LGARB_FINISH_SPW=2
run L_spw4_32k LGARB_FINISH_SPW=4 LGARB_WAVE_FINISH_BELOW=32768
timeout 1200 python bench.py > gpurun_out/c7_bench_full.json 2> gpurun_out/c7_bench_full.err
timeout 1200 python bench.py --workload config3 > gpurun_out/c7_config3.json 2> gpurun_out/c7_config3.err
for f in gpurun_out/c7_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
