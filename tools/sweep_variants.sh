#!/bin/bash
# bench every build/liblgarb_*.so variant (kernel-only line)
for f in build/liblgarb_*.so; do
  v=$(LGARB_LIBRARY=$PWD/$f timeout 300 python bench.py --no-e2e --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g %s' % (d['value'], d['failed_columns']))")
  echo "$f $v"
done
