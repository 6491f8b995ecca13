#!/bin/bash
# usage (on the GPU box): tools/gpu_quick.sh TAG  -> gather/parity tests + default bench into gpurun_out/TAG_*
TAG=${1:-quick}
timeout -s KILL ${QUICK_TEST_TIMEOUT:-240} python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log
timeout -s KILL ${QUICK_BENCH_TIMEOUT:-180} python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<P
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))
for k,v in d['kernels'].items(): print(' ', k, round(v['ms_per_step'],4))
P
cat gpurun_out/${TAG}_pytest.log
