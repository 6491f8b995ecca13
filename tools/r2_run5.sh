#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "gather or volume or hybrid3d or indices or formats" 2>&1 | tail -15 > gpurun_out/r2_run5_pytest.log
JHN_LIB_SUFFIX= timeout -s KILL 300 python bench.py --no-cpu-baseline --no-latency --no-extras > gpurun_out/r2_run5_bench.json 2> gpurun_out/r2_run5_bench.err
python - <<P
import json
for v in ("",):
    try:
        d = json.load(open('gpurun_out/r2_run5_bench%s.json' % v))
        print('variant', v or 'ldgsts', 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'],1), 'cl', round(d['value_f16cl_input']['value'],1))
        for k, v_ in d['kernels'].items(): print('   ', k, v_['launches'], round(v_['ms_per_step'], 4))
    except Exception as e:
        print('bench failed', e); print(open('gpurun_out/r2_run5_bench%s.err' % v).read()[-2000:])
P
cat gpurun_out/r2_run5_pytest.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gather_block -s 1 -c 1 -o gpurun_out/r2_gather_block2 -f python bench.py --no-cpu-baseline --no-latency --no-extras --steps 2 --warmup 1 > gpurun_out/r2_ncu_gather2.log 2>&1
