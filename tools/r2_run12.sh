#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2; do JHN_UPLOAD_MODE=$m timeout 300 python tools/upload_bench.py 2>&1 | tail -3; done | tee gpurun_out/r2_run12_upload.txt
timeout -s KILL 600 python bench.py --no-extras > gpurun_out/r2_run12_bench.json 2> gpurun_out/r2_run12_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_run12_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
P
tail -5 gpurun_out/r2_run12_bench.err
timeout -s KILL 600 python -m pytest tests/test_gpu_reference_dropin.py tests/test_gpu_tensorcore.py -m gpu -q --timeout 600 -k "host or predict_frames" 2>&1 | tail -5
