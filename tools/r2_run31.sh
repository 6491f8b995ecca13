#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensorcore.py tests/test_real_anchor.py -m gpu -q --timeout 120 -x 2>&1 | tail -6
timeout -s KILL 200 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run31_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})" | tee gpurun_out/r2_run31.txt
tail -3 gpurun_out/r2_run31_bench.err
