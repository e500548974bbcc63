#!/bin/bash
# usage: sweep_env.sh "VAR=val VAR2=val" "VAR=val" ...   -- one kernel-only bench line per environment
for envs in "$@"; do
  v=$(env $envs timeout 300 python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g %s %d' % (d['value'], d['failed_columns'], d['gpu_launches']))")
  echo "$envs -> $v"
done
