#!/bin/bash
# round 2, call 17 (TWO GPUs): BASELINE config 4 at its stated size (4 M columns x 7 500 h over 2 GPUs) with the seed-99 parity sample
# of 10 000 columns in the job.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 1200 $T bench.py --gpus 2 --workload config4 > gpurun_out/c17_config4_2gpu.json 2> gpurun_out/c17_config4_2gpu.err; echo "config4 rc $?"
python -c "
import json
d=json.load(open('gpurun_out/c17_config4_2gpu.json'))
print(' value %.4g n_gpus %s failed %s parity %s'%(d['value'], d['n_gpus'], d.get('failed_columns'), json.dumps(d.get('parity'))[:700]))"
grep -n "Error" gpurun_out/c17_config4_2gpu.err | head -5
