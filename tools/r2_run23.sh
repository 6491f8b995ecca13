#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 600 -x 2>&1 | tail -5
run() {
  JHN_LIB_SUFFIX="$1" JHN_NVCC_EXTRA="$2" timeout -s KILL 600 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run23_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), 'conv3', round(d['kernels']['tc_conv3_stacked']['ms_per_step'],4))"
}
{
run "" ""
} | tee gpurun_out/r2_run23.txt
