#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 300 -x 2>&1 | tail -8 > gpurun_out/r2_run6_pytest.log
timeout -s KILL 300 python bench.py --no-cpu-baseline --no-latency --no-extras > gpurun_out/r2_run6_bench.json 2> gpurun_out/r2_run6_bench.err
python - <<P
import json
try:
    d = json.load(open('gpurun_out/r2_run6_bench.json'))
    print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'],1), 'cl', round(d['value_f16cl_input']['value'],1))
    for k, v_ in d['kernels'].items(): print('   ', k, v_['launches'], round(v_['ms_per_step'], 4))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2_run6_bench.err').read()[-2000:])
P
cat gpurun_out/r2_run6_pytest.log
