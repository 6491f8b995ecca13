#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 200 -x -k "host" 2>&1 | tail -3 | tee gpurun_out/r2_run46.txt
timeout -s KILL 300 python tools/e2e_timeline.py hybrid:0.6 32 1 2>&1 | tail -30 > gpurun_out/r2_run46_timeline.txt
timeout -s KILL 500 python tools/e2e_upload_probe.py dma hybrid:0.6/128/48/8/32/1 hybrid:0.6/128/48/8/32/2 hybrid:0.5/128/48/8/32/1 hybrid:0.7/128/48/8/32/1 pull/128/48/8/32/1 dma/128/48/8/8/1 2>&1 | tail -8 | tee -a gpurun_out/r2_run46.txt
