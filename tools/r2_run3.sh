#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -30 > gpurun_out/r2_run3_pytest.log
for v in 0 1; do
JHN_GATHER_V1=$v timeout -s KILL 300 python bench.py --no-cpu-baseline --no-latency --no-extras > gpurun_out/r2_run3_bench_v$v.json 2> gpurun_out/r2_run3_bench_v$v.err
done
python - <<P
import json
for v in (0, 1):
    try:
        d = json.load(open('gpurun_out/r2_run3_bench_v%d.json' % v))
        print('gather v1=%d' % v, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'],1), 'cl', d.get('value_f16cl_input'))
        for k, v_ in d['kernels'].items(): print('   ', k, v_['launches'], round(v_['ms_per_step'], 4))
    except Exception as e:
        print('bench failed', e); print(open('gpurun_out/r2_run3_bench_v%d.err' % v).read()[-2000:])
P
cat gpurun_out/r2_run3_pytest.log
