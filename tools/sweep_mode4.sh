#!/bin/bash
# usage: sweep_mode4.sh "VAR=val VAR2=val" ...   -- one kernel-only bench line (step kernel, mode 4) per environment
for envs in "$@"; do
  v=$(env $envs timeout 300 python bench.py --mode ${MODE:-4} --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g %s %d' % (d['value'], d['failed_columns'], d['gpu_launches']))")
  echo "$envs -> $v"
done
