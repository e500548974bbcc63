#!/bin/bash
# round 2, call 19 (one GPU): fewer, longer search launches in the latency-bound rounds (LGARB_WAVE_LOW_BELOW), three call lengths.
mkdir -p gpurun_out
B="timeout 150 python bench.py --no-cpu-baseline --no-e2e --no-extras"
for low in 0 200000 400000; do
  LGARB_WAVE_LOW_BELOW=$low $B > gpurun_out/c19_low${low}.json 2> gpurun_out/c19_low${low}.err
  LGARB_WAVE_LOW_BELOW=$low $B --steps 1 > gpurun_out/c19_48h_low${low}.json 2> gpurun_out/c19_48h_low${low}.err
  LGARB_WAVE_LOW_BELOW=$low $B --steps 5 > gpurun_out/c19_240h_low${low}.json 2> gpurun_out/c19_240h_low${low}.err
done
for f in gpurun_out/c19_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
