#!/bin/bash
# round 2, first GPU call: the GPU suite with the exact pow + fast search path, then the bench in the new and the old scheduler,
# then what the exact pow and the fast path each cost / bring (variant builds, tools/build_variant.sh), then a debug timeline of
# one 240 h window.  Every bench run here is a plain run (no profiler attached).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt
python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c1_pytest.log
B="python bench.py --no-cpu-baseline"
$B --mode 5 > gpurun_out/c1_bench_m5.json 2> gpurun_out/c1_bench_m5.err
$B --mode 3 --no-e2e > gpurun_out/c1_bench_m3.json 2> gpurun_out/c1_bench_m3.err
$B --mode 5 --steps 20 --no-e2e > gpurun_out/c1_bench_m5_960h.json 2> gpurun_out/c1_bench_m5_960h.err
$B --mode 5 --steps 1 --no-e2e > gpurun_out/c1_bench_m5_48h.json 2> gpurun_out/c1_bench_m5_48h.err
LGARB_LIBRARY=$PWD/build/liblgarb_fastpow.so $B --mode 5 --no-e2e > gpurun_out/c1_bench_m5_fastpow.json 2> gpurun_out/c1_bench_m5_fastpow.err
LGARB_LIBRARY=$PWD/build/liblgarb_fastpow.so $B --mode 3 --no-e2e > gpurun_out/c1_bench_m3_fastpow.json 2> gpurun_out/c1_bench_m3_fastpow.err
LGARB_LIBRARY=$PWD/build/liblgarb_noff.so $B --mode 3 --no-e2e > gpurun_out/c1_bench_m3_noff.json 2> gpurun_out/c1_bench_m3_noff.err
LGARB_LIBRARY=$PWD/build/liblgarb_noff.so $B --mode 5 --no-e2e > gpurun_out/c1_bench_m5_noff.json 2> gpurun_out/c1_bench_m5_noff.err
LGARB_WAVE_DEBUG=1 $B --mode 5 --steps 5 --no-e2e > gpurun_out/c1_dbg_m5.json 2> gpurun_out/c1_dbg_m5.err
tail -5 gpurun_out/c1_pytest.log
for f in gpurun_out/c1_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g  e2e %s  launches %s failed %s'%(d['value'], d['e2e'] and '%.4g'%d['e2e']['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
