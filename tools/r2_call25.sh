#!/bin/bash
# round 2, call 25 (one GPU): the reset kernels folded into the consumers (the last block of a stage kernel clears the queue it
# consumed: half the launches).  GPU suite, two bench lines, then -- only if both are fine -- the ncu launch list of the new build, its
# window profile, and the default line carrying it.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/c25_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c25_pytest.log
tail -4 gpurun_out/c25_pytest.log
grep -q "pytest rc 0" gpurun_out/c25_pytest.log || { echo "GPU suite not green: stopping here"; grep -E "^FAILED|Error" gpurun_out/c25_pytest.log | head -20; exit 1; }
B="timeout 150 python bench.py --no-cpu-baseline --no-e2e --no-extras"
$B > gpurun_out/c25_default.json 2> gpurun_out/c25_default.err
$B --steps 1 > gpurun_out/c25_48h.json 2> gpurun_out/c25_48h.err
for f in gpurun_out/c25_default.json gpurun_out/c25_48h.json; do python -c "
import json; d=json.load(open('$f')); print('$f value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))"; done
bash tools/prof_window.sh; echo "prof_window rc $?"
mv gpurun_out/window_launches.csv gpurun_out/c25_window_launches.csv
python tools/make_window_profile.py gpurun_out/c25_window_launches.csv 1250000 240 > profiles/r2_window_profile.json && cp profiles/r2_window_profile.json gpurun_out/c25_window_profile.json
timeout 600 python bench.py > gpurun_out/c25_bench_full.json 2> gpurun_out/c25_bench_full.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('gpurun_out/c25_bench_full.json')); e=d['e2e']; r=d['roofline']
print('full: value %.4g e2e %.4g launches %s traffic %s hourly %.4g'%(d['value'], e['value'], d['gpu_launches'], r['traffic'], e['hourly_bmi_loop']['value']))"
