#!/bin/bash
mkdir -p gpurun_out
run() {
  JHN_LIB_SUFFIX="$1" JHN_NVCC_EXTRA="$2" timeout -s KILL 600 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run24_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), 'conv3', round(d['kernels']['tc_conv3_stacked']['ms_per_step'],4))"
}
{
run _xA "-DC3_DBG_NO_EPI -DC3_DBG_NO_TMA -DC3_ISSUER_ELECT"
run _xB "-DC3_DBG_NO_EPI -DC3_DBG_NO_TMA"
run _xE "-DC3_ISSUER_ELECT"
} | tee gpurun_out/r2_run24.txt
