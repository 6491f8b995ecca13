#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_real_anchor.py tests/test_gpu_reference_dropin.py -m gpu -q --timeout 600 -x -s 2>&1 | tail -40 > gpurun_out/r2_run8_pytest.log
cat gpurun_out/r2_run8_pytest.log
