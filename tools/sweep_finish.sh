for f in build/liblgarb_F8.so build/liblgarb_F6.so; do for st in 4 8 16; do
  v=$(LGARB_FINISH_LANE_STRIDE=$st LGARB_LIBRARY=$PWD/$f timeout 300 python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g %s' % (d['value'], d['failed_columns']))")
  echo "$f stride $st: $v"
done; done
