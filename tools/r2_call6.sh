#!/bin/bash
# round 2, sixth GPU call: the GPU suite (mode 2 removed, express lane + shared-memory finishing kernel in the scheduling test),
# then what the finishing kernel variants and the express lane bring: calls of 480 h and of 1 460 h (the chunk of the full-year run).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c6_pytest.log
tail -8 gpurun_out/c6_pytest.log
B="timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-extras"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/c6_$name.json 2> gpurun_out/c6_$name.err; }
EXTRA=""
run a_old_ff1      LGARB_EXPRESS=0 LGARB_FINISH_SMEM=0 LGARB_FINISH_FF=1
run b_old_ff0      LGARB_EXPRESS=0 LGARB_FINISH_SMEM=0 LGARB_FINISH_FF=0
run c_smem_ff1     LGARB_EXPRESS=0 LGARB_FINISH_SMEM=1 LGARB_FINISH_FF=1
run d_smem_ff0     LGARB_EXPRESS=0 LGARB_FINISH_SMEM=1 LGARB_FINISH_FF=0
run e_smem_64k     LGARB_EXPRESS=0 LGARB_FINISH_SMEM=1 LGARB_WAVE_FINISH_BELOW=65536
run f_smem_spw16   LGARB_EXPRESS=0 LGARB_FINISH_SMEM=1 LGARB_EXPRESS_SPW=16
run g_express      LGARB_EXPRESS=1
run h_express_x12  LGARB_EXPRESS=1 LGARB_EXPRESS_FACTOR_X4=12 LGARB_EXPRESS_CAP_PERMILLE=80
run i_express_ff   LGARB_EXPRESS=1 LGARB_EXPRESS_FF=1
EXTRA="--hours-per-step 146 --spinup-hours 730"
run L_a_old        LGARB_EXPRESS=0 LGARB_FINISH_SMEM=0
run L_c_smem       LGARB_EXPRESS=0 LGARB_FINISH_SMEM=1
run L_g_express    LGARB_EXPRESS=1
run L_h_express_x12 LGARB_EXPRESS=1 LGARB_EXPRESS_FACTOR_X4=12 LGARB_EXPRESS_CAP_PERMILLE=80
run L_dbg_express  LGARB_EXPRESS=1 LGARB_WAVE_DEBUG=1
run L_dbg_off      LGARB_EXPRESS=0 LGARB_FINISH_SMEM=0 LGARB_WAVE_DEBUG=1
for f in gpurun_out/c6_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(' value %.4g launches %s failed %s'%(d['value'], d['gpu_launches'], d['failed_columns']))
except Exception as ex: print(' FAILED', ex)
"; done
