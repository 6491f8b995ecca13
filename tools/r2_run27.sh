#!/bin/bash
mkdir -p gpurun_out
JHN_LIB_SUFFIX=_d36 JHN_NVCC_EXTRA="-DC3_FIXED_D=36" timeout -s KILL 200 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run27_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})" | tee gpurun_out/r2_run27.txt
