#!/bin/bash
mkdir -p gpurun_out
tools/lds_bench > gpurun_out/r2_lds_bench.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/r2_run2_pytest.log
timeout -s KILL 400 python bench.py > gpurun_out/r2_run2_bench.json 2> gpurun_out/r2_run2_bench.err
python - <<P
import json
try:
    d = json.load(open('gpurun_out/r2_run2_bench.json'))
    print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', d['e2e'], 'cl', d.get('value_f16cl_input'))
    for k, v in d['kernels'].items(): print('   ', k, v['launches'], round(v['ms_per_step'], 4))
    print(json.dumps(d.get('extra'), indent=1)[:3000])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2_run2_bench.err').read()[-3000:])
P
cat gpurun_out/r2_run2_pytest.log
cat gpurun_out/r2_lds_bench.txt
