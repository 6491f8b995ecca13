#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 600 -k "host" 2>&1 | tail -15
timeout -s KILL 600 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run14_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])" | tee gpurun_out/r2_run14.txt
tail -3 gpurun_out/r2_run14_bench.err
