#!/bin/bash
# round 2, fourth GPU call: find the mode that hung call 3 (every step under its own timeout), then the GPU suite with a per-test
# timeout, then the regime switch of the fast search path.
mkdir -p gpurun_out
for m in 3 2 1 0; do timeout 120 python tools/r2_dbg_modes.py $m > gpurun_out/c4_mode$m.log 2>&1; echo "mode $m rc $?" >> gpurun_out/c4_modes.txt; done
LGARB_WAVE_DEBUG=1 timeout 120 python tools/r2_dbg_modes.py 3 > gpurun_out/c4_mode3_dbg.log 2>&1
cat gpurun_out/c4_modes.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=200 --timeout-method=thread > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c4_pytest.log
tail -15 gpurun_out/c4_pytest.log
B="timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extras"
for ff in 0 65536 262144; do
  LGARB_WAVE_FF_BELOW=$ff $B > gpurun_out/c4_bench_ff$ff.json 2> gpurun_out/c4_bench_ff$ff.err
  LGARB_WAVE_FF_BELOW=$ff $B --steps 1 > gpurun_out/c4_bench48_ff$ff.json 2> gpurun_out/c4_bench48_ff$ff.err
done
for f in gpurun_out/c4_bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
