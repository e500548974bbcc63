#!/bin/bash
# round 2, third GPU call: the GPU suite (offender columns, rare paths, path counters), then the regime switch of the fast search
# path: plain walk in every round (LGARB_WAVE_FF_BELOW=0) against the fast path below 65 536 / 262 144 slots in flight.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c3_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e"
for ff in 0 65536 262144; do
  LGARB_WAVE_FF_BELOW=$ff $B > gpurun_out/c3_bench_ff$ff.json 2> gpurun_out/c3_bench_ff$ff.err
  LGARB_WAVE_FF_BELOW=$ff $B --steps 1 > gpurun_out/c3_bench48_ff$ff.json 2> gpurun_out/c3_bench48_ff$ff.err
done
LGARB_WAVE_FF_BELOW=0 LGARB_WAVE_DEBUG=1 $B --steps 5 > gpurun_out/c3_dbg_ff0.json 2> gpurun_out/c3_dbg_ff0.err
LGARB_WAVE_FF_BELOW=65536 LGARB_WAVE_DEBUG=1 $B --steps 5 > gpurun_out/c3_dbg_ff65536.json 2> gpurun_out/c3_dbg_ff65536.err
python bench.py --no-cpu-baseline > gpurun_out/c3_bench_full.json 2> gpurun_out/c3_bench_full.err
tail -5 gpurun_out/c3_pytest.log
for f in gpurun_out/c3_bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g  e2e %s  hourly %s launches %s failed %s'%(d['value'], d['e2e'] and '%.4g'%d['e2e']['value'], d['e2e'] and '%.4g'%d['e2e']['hourly_bmi_loop']['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
