#!/bin/bash
mkdir -p gpurun_out
PROBE_DUMMY=1 timeout -s KILL 300 python tools/coreside_probe.py 2>&1 | tail -12 | tee gpurun_out/r2_run49.txt
