#!/bin/bash
mkdir -p gpurun_out
run() {
  JHN_LIB_SUFFIX="$1" JHN_NVCC_EXTRA="$2" timeout -s KILL 600 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run28_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), 'conv3', round(d['kernels']['tc_conv3_stacked']['ms_per_step'],4))"
}
{
run "" ""
run _ae "-DC3_DBG_NO_EPI"
run _at "-DC3_DBG_NO_TMA"
run _aet "-DC3_DBG_NO_EPI -DC3_DBG_NO_TMA"
} | tee gpurun_out/r2_run28.txt
