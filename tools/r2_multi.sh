#!/bin/bash
# usage: r2_multi.sh N
N=$1
mkdir -p gpurun_out
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_multi_${N}gpu_bench.json 2> gpurun_out/r2_multi_${N}gpu.err
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --workload c5_stress > gpurun_out/r2_multi_${N}gpu_c5_stress.json 2>> gpurun_out/r2_multi_${N}gpu.err
python - <<P
import json
for f in ('gpurun_out/r2_multi_${N}gpu_bench.json','gpurun_out/r2_multi_${N}gpu_c5_stress.json'):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],3), d['e2e']['value'], d['config'].get('workload'))
    except Exception as e: print(f, 'failed', e)
P
tail -5 gpurun_out/r2_multi_${N}gpu.err
