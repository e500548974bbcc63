#!/bin/bash
# round 2, call 18 (EIGHT GPUs): BASELINE config 5 at its stated size (10 M columns x 8 760 h over 8 GPUs) with the seed-99 parity sample
# of 10 000 columns in the job.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/c18_smi.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
timeout 1500 $T bench.py --gpus 8 --workload config5 > gpurun_out/c18_config5_8gpu.json 2> gpurun_out/c18_config5_8gpu.err; echo "config5 rc $?"
python -c "
import json
d=json.load(open('gpurun_out/c18_config5_8gpu.json'))
print(' value %.4g n_gpus %s failed %s parity %s'%(d['value'], d['n_gpus'], d.get('failed_columns'), json.dumps(d.get('parity'))[:700]))"
grep -n "Error" gpurun_out/c18_config5_8gpu.err | head -5
