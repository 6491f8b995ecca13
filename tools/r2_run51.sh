#!/bin/bash
mkdir -p gpurun_out
PROBE_KERNELS=1 timeout -s KILL 300 python tools/coreside_probe.py 128/48/8 128/148/8 32/148/8 2>&1 | tail -9 | tee gpurun_out/r2_run51.txt
timeout -s KILL 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 200 -x -k "host" 2>&1 | tail -3 | tee -a gpurun_out/r2_run51.txt
timeout -s KILL 500 python tools/e2e_upload_probe.py dma hybrid:0.6/128/48/8/32/1 hybrid:0.7/128/48/8/32/1 hybrid:0.8/128/48/8/32/1 pull/128/48/8/32/1 pull/128/148/8/32/1 hybrid:0.7/128/148/8/32/1 hybrid:0.7/128/96/8/32/1 hybrid:0.7/128/96/8/16/1 2>&1 | tail -10 | tee -a gpurun_out/r2_run51.txt
