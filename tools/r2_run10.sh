#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_reference_dropin.py tests/test_gpu_center.py -m gpu -q --timeout 600 -s 2>&1 | tail -150 > gpurun_out/r2_run10_pytest.log
cat gpurun_out/r2_run10_pytest.log
