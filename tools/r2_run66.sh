#!/bin/bash
mkdir -p gpurun_out
P='import json,sys
d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), d["n_gpus"])'
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --no-extras --no-latency 2> gpurun_out/r2_run66_8gpu.err | tee gpurun_out/r2_run66_8gpu_hybrid.json | python -c "$P" hybrid | tee gpurun_out/r2_run66.txt
JHN_E2E_UPLOAD=dma JHN_E2E_AHEAD=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --no-extras --no-latency 2>> gpurun_out/r2_run66_8gpu.err | tee gpurun_out/r2_run66_8gpu_dma.json | python -c "$P" dma | tee -a gpurun_out/r2_run66.txt
