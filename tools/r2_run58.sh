#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 500 python tools/e2e_upload_probe.py dma hybrid:0.5/128/48/8/32/1/2 hybrid:0.6/128/48/8/32/1/2 hybrid:0.7/128/48/8/32/1/2 hybrid:0.5/128/48/8/32/1/3 hybrid:0.6/128/48/8/32/1/3 hybrid:0.7/128/48/8/32/1/3 hybrid:0.5/128/48/8/32/2/3 hybrid:0.6/128/48/8/32/2/3 hybrid:0.7/128/48/8/32/2/3 hybrid:0.6/128/48/8/32/2/4 hybrid:0.8/128/48/8/32/2/3 2>&1 | tail -12 | tee gpurun_out/r2_run58.txt
