#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gather_block -s 1 -c 1 -o gpurun_out/r2_gather_block -f python bench.py --no-cpu-baseline --no-latency --no-extras --steps 2 --warmup 1 > gpurun_out/r2_ncu_gather.log 2>&1
tail -5 gpurun_out/r2_ncu_gather.log
ls -la gpurun_out/*.ncu-rep
