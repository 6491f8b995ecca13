#!/bin/bash
# round 2, GPU call 1: LDS model probe, full GPU test suite, bench at several pass sizes
mkdir -p gpurun_out
tools/lds_bench > gpurun_out/r2_lds_bench.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -25 > gpurun_out/r2_run1_pytest.log
for sb in 32 16 8 4; do
  timeout -s KILL 200 python bench.py --no-cpu-baseline --no-latency --sub-batch $sb > gpurun_out/r2_run1_bench_sb$sb.json 2> gpurun_out/r2_run1_bench_sb$sb.err
done
python - <<P
import json
for sb in (32, 16, 8, 4):
    try:
        d = json.load(open('gpurun_out/r2_run1_bench_sb%d.json' % sb))
        print('sb', sb, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1))
        for k, v in d['kernels'].items(): print('   ', k, v['launches'], round(v['ms_per_step'], 4))
    except Exception as e:
        print('sb', sb, 'failed', e)
P
cat gpurun_out/r2_run1_pytest.log
cat gpurun_out/r2_lds_bench.txt
