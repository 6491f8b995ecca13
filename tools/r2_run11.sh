#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 > gpurun_out/r2_run11_pytest.log
cat gpurun_out/r2_run11_pytest.log | cut -c1-250
timeout -s KILL 600 python bench.py > gpurun_out/r2_run11_bench.json 2> gpurun_out/r2_run11_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_run11_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
P
tail -5 gpurun_out/r2_run11_bench.err
