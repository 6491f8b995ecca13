#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/upload_bench.py 2>&1 | tail -1 | tee gpurun_out/r2_run17.txt
timeout -s KILL 600 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run17_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])" | tee -a gpurun_out/r2_run17.txt
tail -3 gpurun_out/r2_run17_bench.err
