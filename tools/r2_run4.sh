#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "gather or volume or hybrid3d or indices or formats" 2>&1 | tail -15 > gpurun_out/r2_run4_pytest.log
JHN_LIB_SUFFIX= timeout -s KILL 300 python bench.py --no-cpu-baseline --no-latency --no-extras > gpurun_out/r2_run4_bench.json 2> gpurun_out/r2_run4_bench.err
JHN_LIB_SUFFIX=_r80 JHN_NVCC_EXTRA="-DGB_PROD_REGS=80" timeout -s KILL 300 python bench.py --no-cpu-baseline --no-latency --no-extras > gpurun_out/r2_run4_bench_r80.json 2> gpurun_out/r2_run4_bench_r80.err
python - <<P
import json
for v in ("", "_r80"):
    try:
        d = json.load(open('gpurun_out/r2_run4_bench%s.json' % v))
        print('variant', v or 'r72', 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'],1), 'cl', round(d['value_f16cl_input']['value'],1))
        for k, v_ in d['kernels'].items(): print('   ', k, v_['launches'], round(v_['ms_per_step'], 4))
    except Exception as e:
        print('bench failed', e); print(open('gpurun_out/r2_run4_bench%s.err' % v).read()[-2000:])
P
cat gpurun_out/r2_run4_pytest.log
