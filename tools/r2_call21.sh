#!/bin/bash
# round 2, call 21 (one GPU): the live prefix of the front arrays loaded chunk by chunk across ALL arrays (5 loads in flight per thread
# instead of 1), and a variant that requests the next slot's record into L2 before writing the current one back.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/c21_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c21_pytest.log
tail -4 gpurun_out/c21_pytest.log
grep -q "pytest rc 0" gpurun_out/c21_pytest.log || { echo "GPU suite not green: stopping here"; grep -E "^FAILED|Error" gpurun_out/c21_pytest.log | head -20; exit 1; }
B="timeout 150 python bench.py --no-cpu-baseline --no-e2e --no-extras"
$B > gpurun_out/c21_default.json 2> gpurun_out/c21_default.err
$B > gpurun_out/c21_default_b.json 2> gpurun_out/c21_default_b.err
LGARB_LIBRARY=$PWD/build/liblgarb_prefetch.so $B > gpurun_out/c21_prefetch.json 2> gpurun_out/c21_prefetch.err
LGARB_LIBRARY=$PWD/build/liblgarb_prefetch.so $B > gpurun_out/c21_prefetch_b.json 2> gpurun_out/c21_prefetch_b.err
$B --hours-per-step 146 --spinup-hours 730 > gpurun_out/c21_L_default.json 2> gpurun_out/c21_L_default.err
LGARB_LIBRARY=$PWD/build/liblgarb_prefetch.so $B --hours-per-step 146 --spinup-hours 730 > gpurun_out/c21_L_prefetch.json 2> gpurun_out/c21_L_prefetch.err
for f in gpurun_out/c21_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
