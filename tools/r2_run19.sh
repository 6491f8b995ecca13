#!/bin/bash
mkdir -p gpurun_out
run() { # suffix flags
  JHN_LIB_SUFFIX="$1" JHN_NVCC_EXTRA="$2" timeout -s KILL 600 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run19_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), 'norm', round(d['kernels']['tc_norm_act_kernel']['ms_per_step'],4))"
}
{
run "" ""
run _na "-DNORM_UNROLL_V=4 -DNORM_MINB=4"
run _nb "-DNORM_UNROLL_V=2 -DNORM_MINB=6"
run _nc "-DNORM_UNROLL_V=6 -DNORM_MINB=3"
run _nd "-DNORM_UNROLL_V=4 -DNORM_MINB=4 -DNORM_ZB_V=9"
} | tee gpurun_out/r2_run19.txt
tail -3 gpurun_out/r2_run19_bench.err
