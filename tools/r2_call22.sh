#!/bin/bash
# round 2, call 22 (EIGHT GPUs): the default line on 8 ranks (weak scaling, NCCL gather block), then BASELINE config 4 over 4 of them.
mkdir -p gpurun_out
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522"
timeout 600 $T8 bench.py --gpus 8 --no-cpu-baseline --no-extras > gpurun_out/c22_bench_8gpu.json 2> gpurun_out/c22_bench_8gpu.err; echo "bench 8 rc $?"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523"
timeout 600 $T4 bench.py --gpus 4 --workload config4 > gpurun_out/c22_config4_4gpu.json 2> gpurun_out/c22_config4_4gpu.err; echo "config4 x4 rc $?"
for f in gpurun_out/c22_bench_8gpu.json gpurun_out/c22_config4_4gpu.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); e=d.get('e2e') or {}
    print(' value %.4g n_gpus %s failed %s e2e %s gather %s parity %s'%(d['value'], d['n_gpus'], d.get('failed_columns'), e.get('value'), json.dumps(d.get('gather'))[:200], json.dumps(d.get('parity'))[:330]))
except Exception as ex: print(' FAILED', ex)
"; done
