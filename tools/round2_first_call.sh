#!/bin/bash
# First GPU call of the next round: the variants written after round 1's GPU budget was spent.
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh > gpurun_out/round2_first_call.log 2>&1'
# 1. bit-identity of the shared-memory finishing kernel (lgar_finish_smem.cu) with the default scheduler;
# 2. what it is worth: bench value (device-resident, 480 h call) against the hand-over threshold -- the point of a faster
#    finishing kernel is to hand over EARLIER and skip the latency-bound drop-out rounds.
set -u
cd "$(dirname "$0")/.."
LGARB_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_scale_parity.py -x -q -m gpu -k experimental || echo "EXPERIMENTAL VARIANTS FAILED"
run() {  # label, env...
  local label=$1; shift
  local v
  v=$(env "$@" timeout 300 python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g failed=%s' % (d['value'], d['failed_columns']))")
  echo "$label: $v"
}
run "default (local records, stride 4, below 16384)"
for below in 16384 32768 65536 131072 262144; do
  run "local  stride 4 below $below" LGARB_WAVE_FINISH_BELOW=$below
  for st in 2 4 8; do
    run "smem   stride $st below $below" LGARB_FINISH_SMEM=1 LGARB_FINISH_LANE_STRIDE=$st LGARB_WAVE_FINISH_BELOW=$below
  done
done
# 3. the converged shared-memory kernel (LGARB_FINISH_SMEM=2): as a finishing kernel, and running the WHOLE window on chip
#    (hand-over threshold above the column count) -- the measurement behind DESIGN.md section 10 item 1
for below in 16384 65536 262144 2000000; do
  for q in 0 -32 -512; do
    run "vote-smem quantum ${q#-} below $below" LGARB_FINISH_SMEM=2 LGARB_FINISH_LANE_STRIDE=$q LGARB_WAVE_FINISH_BELOW=$below
  done
done
# 4. one `ncu --set full` capture each of the default finishing kernel and of the two shared-memory ones (timed window 8 of
#    the profiling command of tools/prof_window.sh; the finishing kernel is the last launch of a window)
B="python bench.py --mode 3 --no-e2e --no-cpu-baseline --spinup-hours 240 --steps 5"
export LGARB_NCU_WINDOW=8 LGARB_NCU_ROUND=-2
for v in "local" "smem1 LGARB_FINISH_SMEM=1" "smem2 LGARB_FINISH_SMEM=2"; do
  set -- $v; name=$1; shift
  env "$@" timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name regex:lgar_wave_finish \
      -o gpurun_out/r2_finish_$name $B > gpurun_out/r2_finish_$name.log 2>&1 || echo "ncu capture $name failed"
done
