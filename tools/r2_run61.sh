#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_abi.py -m gpu -q --timeout 200 -x 2>&1 | tail -4 | tee gpurun_out/r2_run61.txt
timeout -s KILL 300 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run61_bench.err | tee gpurun_out/r2_run61_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], round(d['value_f16cl_input']['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})" | tee -a gpurun_out/r2_run61.txt
