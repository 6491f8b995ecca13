#!/bin/bash
mkdir -p gpurun_out
PULL_FROM_DEVICE=1 PULL_REPS=400 PROBE_KERNELS=1 timeout -s KILL 300 python tools/coreside_probe.py 128/48/8 32/148/8 2>&1 | tail -8 | tee gpurun_out/r2_run47.txt
