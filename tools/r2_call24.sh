#!/bin/bash
# round 2, call 24 (one GPU): resident blocks per SM of the SEARCH stage kernel (variant builds: 6 / 8 blocks = 80 / 64 registers).
mkdir -p gpurun_out
B="timeout 150 python bench.py --no-cpu-baseline --no-e2e --no-extras"
$B > gpurun_out/c24_default.json 2> gpurun_out/c24_default.err
LGARB_LIBRARY=$PWD/build/liblgarb_s6.so $B > gpurun_out/c24_s6.json 2> gpurun_out/c24_s6.err
LGARB_LIBRARY=$PWD/build/liblgarb_s8.so $B > gpurun_out/c24_s8.json 2> gpurun_out/c24_s8.err
$B > gpurun_out/c24_default_b.json 2> gpurun_out/c24_default_b.err
LGARB_LIBRARY=$PWD/build/liblgarb_s6.so $B > gpurun_out/c24_s6_b.json 2> gpurun_out/c24_s6_b.err
for f in gpurun_out/c24_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
