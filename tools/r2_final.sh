#!/bin/bash
# round 2 evidence run: whole GPU suite, default bench (the driver's command), reference arm, ncu launch list and one full capture
TAG=${1:-r2_final}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 | cut -c1-300 > gpurun_out/${TAG}_pytest.log
timeout -s KILL 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout -s KILL 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --no-cpu-baseline --no-latency --no-extras --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_b.log 2>&1
# one full forward (27 launches) of the resident path: skip the weight packing + warm-up launches
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"coarse_project|relayout_pixel|gather_stream|tc_conv|tc_norm|tc_head|tc_zero|centroid_finalize" -s 29 -c 27 \
    -o gpurun_out/${TAG}_full -f python bench.py --no-cpu-baseline --no-latency --no-extras --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_f.log 2>&1
python - <<P
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', d['e2e'])
print('roofline', d['roofline']); print('cpu_baseline', d['cpu_baseline'])
for k,v in d['kernels'].items(): print(' ', k, v['launches'], round(v['ms_per_step'],4))
print(open('gpurun_out/${TAG}_bench_reference.json').read()[:600])
P
cat gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_ncu_f.log | cut -c1-200
