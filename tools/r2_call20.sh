#!/bin/bash
# round 2, call 20 -- the evidence of the FINAL build: GPU suite, ncu launch list of one timed window (-> profiles/r2_window_profile.json,
# keyed by the hash of the sources), the default bench line with every extra (now carrying the measured traffic), the reference arm,
# BASELINE config 3 at size, the hourly loop, the stress variant, ncu --set full of one bulk round.  Every ncu capture follows a plain
# run of the same command that exited 0.
mkdir -p gpurun_out
sha256sum lgar_c_b200/liblgarb.so > gpurun_out/c20_lib_sha256.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c20_pytest.log
tail -4 gpurun_out/c20_pytest.log
grep -q "pytest rc 0" gpurun_out/c20_pytest.log || { echo "GPU suite not green: stopping here"; grep -E "^FAILED|Error" gpurun_out/c20_pytest.log | head -20; exit 1; }
bash tools/prof_window.sh; echo "prof_window rc $?"
mv gpurun_out/window_launches.csv gpurun_out/c20_window_launches.csv
python tools/make_window_profile.py gpurun_out/c20_window_launches.csv 1250000 240 > profiles/r2_window_profile.json && cp profiles/r2_window_profile.json gpurun_out/c20_window_profile.json
timeout 900 python bench.py > gpurun_out/c20_bench_full.json 2> gpurun_out/c20_bench_full.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference > gpurun_out/c20_bench_reference.json 2> gpurun_out/c20_bench_reference.err; echo "reference arm rc $?"
timeout 900 python bench.py --workload config3 > gpurun_out/c20_config3.json 2> gpurun_out/c20_config3.err; echo "config3 rc $?"
timeout 150 python tools/hourly_profile.py > gpurun_out/c20_hourly.json 2> gpurun_out/c20_hourly.err
timeout 300 python tools/bench_stress.py --columns 1000000 --hours 12 --cpu > gpurun_out/c20_stress_1m.json 2> gpurun_out/c20_stress_1m.err
bash tools/prof_round.sh 1 full; echo "prof_round rc $?"
if [ -f gpurun_out/round1_full.ncu-rep ]; then
  ncu -i gpurun_out/round1_full.ncu-rep --page raw --csv > gpurun_out/c20_round1_full_raw.csv 2> /dev/null
  python tools/ncu_rep_table.py gpurun_out/round1_full.ncu-rep > gpurun_out/c20_round1_table.md 2> /dev/null
  ncu -i gpurun_out/round1_full.ncu-rep --page source --csv --kernel-name regex:lgar_wave_kernel > gpurun_out/c20_round1_source.csv 2> /dev/null
  [ $(stat -c %s gpurun_out/round1_full.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/round1_full.ncu-rep
  gzip -f gpurun_out/c20_round1_source.csv
fi
cat gpurun_out/c20_hourly.json gpurun_out/c20_stress_1m.json | cut -c1-300
for f in gpurun_out/c20_bench_full.json gpurun_out/c20_bench_reference.json gpurun_out/c20_config3.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); e=d.get('e2e') or {}; r=d.get('roofline') or {}
    print(' value %.4g launches %s failed %s e2e %s traffic %s'%(d['value'], d.get('gpu_launches'), d.get('failed_columns'), e.get('value'), r.get('traffic')))
except Exception as ex: print(' FAILED', ex)
"; done
