#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -80 > gpurun_out/r2_run9_pytest.log
cat gpurun_out/r2_run9_pytest.log
timeout -s KILL 600 python bench.py > gpurun_out/r2_run9_bench.json 2> gpurun_out/r2_run9_bench.err
tail -c 600 gpurun_out/r2_run9_bench.json
