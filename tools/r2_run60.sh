#!/bin/bash
mkdir -p gpurun_out
P='import json,sys
d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items() if k in ("tc_conv3_stacked","tc_norm_act_kernel")})'
echo "maxnreg 120 (shipped)" | tee gpurun_out/r2_run60.txt
timeout -s KILL 200 python bench.py --no-extras --no-latency 2> gpurun_out/r2_run60_bench.err | python -c "$P" | tee -a gpurun_out/r2_run60.txt
echo "maxnreg 128 (rebuilt on the box)" | tee -a gpurun_out/r2_run60.txt
export JHN_NVCC_EXTRA="-DC3_MAXNREG_480=128"
timeout -s KILL 600 python jarvis-hybridnet_b200/build.py > gpurun_out/r2_run60_build.log 2>&1
timeout -s KILL 200 python bench.py --no-extras --no-latency 2>> gpurun_out/r2_run60_bench.err | python -c "$P" | tee -a gpurun_out/r2_run60.txt
timeout -s KILL 200 python bench.py --no-extras --no-latency 2>> gpurun_out/r2_run60_bench.err | python -c "$P" | tee -a gpurun_out/r2_run60.txt
