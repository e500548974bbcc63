#!/bin/bash
# round 2, call 14 (one GPU): the GEFF stage (numeric calc_Geff calls of lg_dzdt_calc as a stage of their own): GPU suite, the stress
# variant of config 3 at 1 M columns with its timeline, the default line as a check that closed-form configurations are where they were.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c14_pytest.log
tail -4 gpurun_out/c14_pytest.log
grep -q "pytest rc 0" gpurun_out/c14_pytest.log || { echo "GPU suite not green: stopping here"; grep -E "^FAILED|Error" gpurun_out/c14_pytest.log | head -20; exit 1; }
timeout 300 python tools/bench_stress.py --columns 1000000 --hours 12 > gpurun_out/c14_stress_1m.json 2> gpurun_out/c14_stress_1m.err
cat gpurun_out/c14_stress_1m.json
LGARB_WAVE_DEBUG=1 timeout 300 python tools/bench_stress.py --columns 1000000 --hours 12 > gpurun_out/c14_stress_dbg.json 2> gpurun_out/c14_stress_dbg.err
grep "wave\]" gpurun_out/c14_stress_dbg.err | tail -12 | cut -c1-300
timeout 150 python bench.py --no-cpu-baseline --no-e2e --no-extras > gpurun_out/c14_default.json 2> gpurun_out/c14_default.err
python -c "
import json; d=json.load(open('gpurun_out/c14_default.json')); print('default value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))"
