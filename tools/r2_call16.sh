#!/bin/bash
# round 2, call 16 (one GPU): every stage moves only the parts of the lane record it reads or writes (FRONT: fronts + 128 bytes of
# step context + state_previous in + search record; HEAD: no epilogue state): GPU suite, then the default line and a call of 1 460 h.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/c16_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c16_pytest.log
tail -4 gpurun_out/c16_pytest.log
grep -q "pytest rc 0" gpurun_out/c16_pytest.log || { echo "GPU suite not green: stopping here"; grep -E "^FAILED|Error" gpurun_out/c16_pytest.log | head -20; exit 1; }
B="timeout 150 python bench.py --no-cpu-baseline --no-e2e --no-extras"
$B > gpurun_out/c16_default.json 2> gpurun_out/c16_default.err
$B > gpurun_out/c16_default_b.json 2> gpurun_out/c16_default_b.err
LGARB_WAVE_DEBUG=1 $B > gpurun_out/c16_default_dbg.json 2> gpurun_out/c16_default_dbg.err
$B --hours-per-step 146 --spinup-hours 730 > gpurun_out/c16_L_default.json 2> gpurun_out/c16_L_default.err
for f in gpurun_out/c16_*default*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
