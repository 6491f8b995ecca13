#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 300 -x 2>&1 | tail -15 > gpurun_out/r2_run7_pytest.log
timeout -s KILL 300 python bench.py --no-cpu-baseline --no-latency --no-extras > gpurun_out/r2_run7_bench.json 2> gpurun_out/r2_run7_bench.err
JHN_NORM_FUSE=0 timeout -s KILL 300 python bench.py --no-cpu-baseline --no-latency --no-extras > gpurun_out/r2_run7_bench_nofuse.json 2> gpurun_out/r2_run7_bench_nofuse.err
python - <<P
import json
for f in ('r2_run7_bench', 'r2_run7_bench_nofuse'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'],1), 'cl', round(d['value_f16cl_input']['value'],1))
        for k, v_ in d['kernels'].items(): print('   ', k, v_['launches'], round(v_['ms_per_step'], 4))
    except Exception as e:
        print(f, 'bench failed', e); print(open('gpurun_out/%s.err' % f).read()[-2000:])
P
cat gpurun_out/r2_run7_pytest.log
