#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 500 python tools/e2e_upload_probe.py dma hybrid:0.6/128/48/8/32/1 hybrid:0.6/128/48/8/32/1 2>&1 | tail -3 | tee gpurun_out/r2_run56.txt
timeout -s KILL 300 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run56_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['steps'])" | tee -a gpurun_out/r2_run56.txt
timeout -s KILL 300 python bench.py --no-extras --no-latency --steps 30 2> gpurun_out/r2_run56_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['steps'])" | tee -a gpurun_out/r2_run56.txt
