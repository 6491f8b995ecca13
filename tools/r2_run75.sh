#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 180 -k "tiny or small" > gpurun_out/r2_run75_racecheck_fp32.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Error: Race" gpurun_out/r2_run75_racecheck_fp32.log | sed -E 's/\+0x[0-9a-f]+//g' | sort | uniq -c | tail -6
timeout -s KILL 540 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests -m gpu -q --timeout 500 -x --deselect tests/test_gpu_ingest.py::test_pull_heatmap_spans_moves_exactly_the_row_spans > gpurun_out/r2_run75_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_run75_memcheck.log | tail -3
