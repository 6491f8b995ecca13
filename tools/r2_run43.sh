#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_run43.txt
for h in 0 1 2 3; do
JHN_C3_L2HINT=$h timeout -s KILL 200 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run43_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('hint $h', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(d['value_f16cl_input']['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})" | tee -a gpurun_out/r2_run43.txt
done
