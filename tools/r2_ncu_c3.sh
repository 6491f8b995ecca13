#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv3 -s 7 -c 1 -o gpurun_out/r2_c3 -f python bench.py --no-cpu-baseline --no-latency --no-extras --steps 2 --warmup 1 > gpurun_out/r2_ncu_c3.log 2>&1
tail -3 gpurun_out/r2_ncu_c3.log | cut -c1-300
ls -la gpurun_out/r2_c3.ncu-rep
