#!/bin/bash
mkdir -p gpurun_out
JHN_LIB_SUFFIX=_xB JHN_NVCC_EXTRA="-DC3_DBG_NO_EPI -DC3_DBG_NO_TMA" timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv3 -s 7 -c 1 -o gpurun_out/r2_c3b -f python bench.py --no-cpu-baseline --no-latency --no-extras --steps 2 --warmup 1 > gpurun_out/r2_ncu_c3b.log 2>&1
tail -2 gpurun_out/r2_ncu_c3b.log | cut -c1-200
