#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 5 -c 5 -o gpurun_out/r2_tc_conv -f python bench.py --no-cpu-baseline --no-latency --no-extras --steps 2 --warmup 1 > gpurun_out/r2_ncu_conv.log 2>&1
tail -3 gpurun_out/r2_ncu_conv.log | cut -c1-300
ls -la gpurun_out/r2_tc_conv.ncu-rep
