#!/bin/bash
# round 2, call 13 (one GPU): the GPU suite of the build with the multi-device entries, then resident blocks per SM of HEAD / FRONT /
# TAIL now that their warps are converged (variant builds with __launch_bounds__ minimum blocks 4 / 6 / 8 = 128 / 80 / 64 registers).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c13_pytest.log
tail -4 gpurun_out/c13_pytest.log
grep -q "pytest rc 0" gpurun_out/c13_pytest.log || { echo "GPU suite not green: stopping here"; grep -E "^FAILED|Error" gpurun_out/c13_pytest.log | head -20; exit 1; }
B="timeout 150 python bench.py --no-cpu-baseline --no-e2e --no-extras"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/c13_$name.json 2> gpurun_out/c13_$name.err; }
EXTRA=""
run default X=1
run mb4 LGARB_LIBRARY=$PWD/build/liblgarb_mb4.so
run mb6 LGARB_LIBRARY=$PWD/build/liblgarb_mb6.so
run mb8 LGARB_LIBRARY=$PWD/build/liblgarb_mb8.so
run mb8f LGARB_LIBRARY=$PWD/build/liblgarb_mb8f.so
EXTRA="--hours-per-step 146 --spinup-hours 730"
run L_default X=1
run L_mb6 LGARB_LIBRARY=$PWD/build/liblgarb_mb6.so
run L_mb8 LGARB_LIBRARY=$PWD/build/liblgarb_mb8.so
for f in gpurun_out/c13_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
