#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 500 python tools/e2e_upload_probe.py dma hybrid-box:0.6/128/48/8/32/1 hybrid:0.6/128/48/8/32/1 hybrid:0.7/128/48/8/32/1 hybrid:0.8/128/48/8/32/1 hybrid-box:0.6/128/48/8/32/1 hybrid:0.6/128/48/8/32/1 hybrid:0.7/128/48/8/32/1 hybrid:0.8/128/96/8/32/1 hybrid-box:0.6/128/96/8/32/1 dma 2>&1 | tail -12 | tee gpurun_out/r2_run54.txt
