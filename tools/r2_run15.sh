#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/upload_bench.py 2>&1 | tail -1 | tee gpurun_out/r2_run15_upload.txt
