#!/bin/bash
mkdir -p gpurun_out
for c in 4 8 16 32; do JHN_E2E_CHUNK=$c timeout -s KILL 200 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run29_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print($c, round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3))"; done | tee gpurun_out/r2_run29.txt
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12 | cut -c1-300 | tee gpurun_out/r2_run29_pytest.log
