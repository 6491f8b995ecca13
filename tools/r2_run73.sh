#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 420 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_center.py tests/test_gpu_ingest.py -m gpu -q --timeout 400 -k "hybrid3d_bf16_end_to_end or fused_head_argmax or center_locate or crop or efftrack or pull_heatmap_spans or (tc_layer and 5-12-1)" > gpurun_out/r2_run73_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Error: Race" gpurun_out/r2_run73_racecheck.log | sort | uniq -c | tail -6
