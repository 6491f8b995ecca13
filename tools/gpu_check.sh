#!/bin/bash
# quick GPU check of a kernel change: the tensor-core / parity tests, then the per-kernel times of one bench run
mkdir -p gpurun_out
TAG=${1:-check}
timeout -s KILL 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -x 2>&1 | tail -4 | tee gpurun_out/${TAG}.txt
timeout -s KILL 300 python bench.py --no-extras --no-latency 2> gpurun_out/${TAG}_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(d['value_f16cl_input']['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})" | tee -a gpurun_out/${TAG}.txt
