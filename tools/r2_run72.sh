#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/racecheck_upload_probe.py 2>&1 | tail -5 | tee gpurun_out/r2_run72.txt
PROBE_KERNELS=1 timeout -s KILL 300 python tools/coreside_probe.py 128/48/8 1/148/8 2>&1 | tail -6 | tee -a gpurun_out/r2_run72.txt
