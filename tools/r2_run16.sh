#!/bin/bash
mkdir -p gpurun_out
for r in 0 3072; do JHN_SMEM_RESERVE=$r timeout -s KILL 600 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run16_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print($r, d['value'], d['ms_per_step'], d['e2e'], {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; done | tee gpurun_out/r2_run16.txt
tail -3 gpurun_out/r2_run16_bench.err
