#!/bin/bash
mkdir -p gpurun_out
for l in 1 2 4 8; do JHN_UPLOAD_LANES=$l timeout 300 python tools/upload_bench.py 2>&1 | tail -1; done | tee gpurun_out/r2_run13_upload.txt
for l in 2 4 8; do JHN_UPLOAD_LANES=$l timeout -s KILL 600 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run13_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print($l, d['value'], d['ms_per_step'], d['e2e'])"; done | tee -a gpurun_out/r2_run13_upload.txt
