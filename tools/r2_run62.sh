#!/bin/bash
mkdir -p gpurun_out
PROBE_OVERLAP=0 timeout -s KILL 300 python tools/coreside_probe.py 128/48/8 2>&1 | tail -2 | tee gpurun_out/r2_run62.txt
PROBE_OVERLAP=1 timeout -s KILL 300 python tools/coreside_probe.py 128/48/8 2>&1 | tail -2 | tee -a gpurun_out/r2_run62.txt
timeout -s KILL 500 python tools/e2e_upload_probe.py dma hybrid:0.6/128/48/8/32/2/3 hybrid:0.6/128/48/8/32/2/3 hybrid:0.5/128/48/8/32/2/3 2>&1 | tail -4 | tee -a gpurun_out/r2_run62.txt
for i in 1 2; do timeout -s KILL 300 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run62_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['e2e']['value'], d['e2e']['ms_per_step'])" | tee -a gpurun_out/r2_run62.txt; done
