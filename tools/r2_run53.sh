#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 200 -x -k "host" 2>&1 | tail -12 | tee gpurun_out/r2_run53.txt
timeout -s KILL 500 python tools/e2e_upload_probe.py dma hybrid:0.6/128/48/8/32/1 hybrid:0.7/128/48/8/32/1 hybrid:0.8/128/48/8/32/1 hybrid:1.0/128/48/8/32/1 hybrid:0.7/128/96/8/32/1 hybrid:0.8/128/96/8/32/1 hybrid:0.8/32/148/8/32/1 2>&1 | tail -10 | tee -a gpurun_out/r2_run53.txt
