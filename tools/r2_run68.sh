#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 400 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 800 -x -k "test_heatmap_formats_agree and tiny_s0" > gpurun_out/r2_run68_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|Race reported|passed|failed" gpurun_out/r2_run68_racecheck.log | sort | uniq -c | head -12
bash tools/gpu_check.sh r2_run68
