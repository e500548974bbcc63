#!/bin/bash
# round 2, second GPU call: the GPU suite, the bench with the single-evaluation-site fast search path, a debug timeline of a
# 240 h window, the full-duration parity sampler at a moderate size.  Plain runs only (no profiler attached).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c2_pytest.log
B="python bench.py --no-cpu-baseline"
$B > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
$B --steps 20 --no-e2e > gpurun_out/c2_bench_960h.json 2> gpurun_out/c2_bench_960h.err
$B --steps 1 --no-e2e > gpurun_out/c2_bench_48h.json 2> gpurun_out/c2_bench_48h.err
LGARB_WAVE_DEBUG=1 $B --steps 5 --no-e2e > gpurun_out/c2_dbg.json 2> gpurun_out/c2_dbg.err
python tools/parity_sampling_full.py --configs 5,4,3 --sample 2000 --batch 100000 --steps 2400 --out gpurun_out/c2_parity_2400h.json > gpurun_out/c2_parity.log 2>&1
tail -5 gpurun_out/c2_pytest.log
tail -4 gpurun_out/c2_parity.log
for f in gpurun_out/c2_bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g  e2e %s  hourly %s launches %s failed %s'%(d['value'], d['e2e'] and '%.4g'%d['e2e']['value'], d['e2e'] and '%.4g'%d['e2e']['hourly_bmi_loop']['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
