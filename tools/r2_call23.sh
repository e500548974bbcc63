#!/bin/bash
# round 2, call 23 (one GPU): what the driver runs at the end of the round, on the committed state: smoke(), the GPU suite, the bench
# with the driver's flags of round 1 (--steps 20 --warmup 3) for both arms.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c23_smoke.txt 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/c23_smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c23_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/c23_pytest.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/c23_bench_reference_steps20.json 2> gpurun_out/c23_bench_reference_steps20.err ) 2>&1 | grep real
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/c23_bench_steps20.json 2> gpurun_out/c23_bench_steps20.err ) 2>&1 | grep real
python -c "
import json
for f in ('gpurun_out/c23_bench_reference_steps20.json','gpurun_out/c23_bench_steps20.json'):
    d=json.load(open(f)); e=d.get('e2e') or {}; r=d.get('roofline') or {}
    print(f, ' value %.4g e2e %.4g launches %s traffic %s frac %s'%(d['value'], e.get('value',0), d.get('gpu_launches'), r.get('traffic'), r.get('frac')))"
