#!/bin/bash
# round 2, fifth GPU call (after the container was re-created): which scheduler mode hangs (each under its own timeout), the GPU
# suite with a per-test timeout, the default bench line with every extra, the debug timeline of one 480 h call.
mkdir -p gpurun_out
for m in 3 2 1 0; do timeout 90 python tools/r2_dbg_modes.py $m > gpurun_out/c5_mode$m.log 2>&1; echo "mode $m rc $?" >> gpurun_out/c5_modes.txt; done
cat gpurun_out/c5_modes.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout=240 --timeout-method=thread > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c5_pytest.log
tail -15 gpurun_out/c5_pytest.log
B="timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extras"
for ff in 0 65536; do
  LGARB_WAVE_FF_BELOW=$ff $B > gpurun_out/c5_bench_ff$ff.json 2> gpurun_out/c5_bench_ff$ff.err
  LGARB_WAVE_FF_BELOW=$ff $B --steps 1 > gpurun_out/c5_bench48_ff$ff.json 2> gpurun_out/c5_bench48_ff$ff.err
done
LGARB_WAVE_DEBUG=1 $B --steps 10 > gpurun_out/c5_dbg.json 2> gpurun_out/c5_dbg.err
timeout 900 python bench.py > gpurun_out/c5_bench_full.json 2> gpurun_out/c5_bench_full.err
for f in gpurun_out/c5_bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
