#!/bin/bash
# round 2, call 15 (TWO GPUs): the multi-device group library (NCCL gather), the standalone driver over two devices against one,
# BASELINE config 4 at its stated size (4 M columns x 7 500 h over 2 GPUs) with the seed-99 parity sample in the job, the default
# line on two ranks (NCCL gather block).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/c15_smi.txt
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py::test_group_of_one_device -m gpu -q -s --timeout=300 --timeout-method=thread > gpurun_out/c15_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c15_pytest.log
tail -5 gpurun_out/c15_pytest.log
CFG=data/configs/config_lasam_Phillipsburg_with_reservoir.txt
mkdir -p /tmp/one /tmp/two
python - <<'PY' > gpurun_out/c15_cfg.txt
import sys; sys.path.insert(0, '.')
from oracle.oracle_py import absolutize_config, DATA
print(absolutize_config(DATA / "configs" / "config_lasam_Phillipsburg_with_reservoir.txt", out_dir="/tmp"))
PY
ACFG=$(cat gpurun_out/c15_cfg.txt)
timeout 300 ./lgar_c_b200/lasam_b200_standalone $ACFG --columns 20001 --shift 7 --write-columns 0,10000,20000 --out-dir /tmp/one > gpurun_out/c15_standalone_one.txt 2>&1
timeout 300 ./lgar_c_b200/lasam_b200_standalone $ACFG --columns 20001 --shift 7 --write-columns 0,10000,20000 --out-dir /tmp/two --devices 0,1 > gpurun_out/c15_standalone_two.txt 2>&1
for c in 0 10000 20000; do cmp /tmp/one/data_variables_c$c.csv /tmp/two/data_variables_c$c.csv && echo "column $c: files of the two-device run identical to the one-device run" >> gpurun_out/c15_standalone_two.txt; done
tail -6 gpurun_out/c15_standalone_two.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 1200 $T bench.py --gpus 2 --workload config4 > gpurun_out/c15_config4_2gpu.json 2> gpurun_out/c15_config4_2gpu.err; echo "config4 rc $?"
timeout 900 $T bench.py --gpus 2 --no-cpu-baseline > gpurun_out/c15_bench_2gpu.json 2> gpurun_out/c15_bench_2gpu.err; echo "bench 2 rc $?"
for f in gpurun_out/c15_config4_2gpu.json gpurun_out/c15_bench_2gpu.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); e=d.get('e2e') or {}
    print(' value %.4g n_gpus %s failed %s e2e %s parity %s gather %s'%(d['value'], d['n_gpus'], d.get('failed_columns'), e.get('value'), json.dumps(d.get('parity'))[:400], json.dumps(d.get('gather'))[:300]))
except Exception as ex: print(' FAILED', ex)
"; done
tail -3 gpurun_out/c15_config4_2gpu.err
