#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2/c39_tests.txt 2>&1
tail -4 gpurun_out/r2/c39_tests.txt
timeout 300 python scripts/stress_filter.py 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/c39_bench.json 2> gpurun_out/r2/c39_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/c39_bench.json') if l.startswith('{')][-1])
print('cfg2', d['value'], d['ms_per_step'], d['stages_us'], 'e2e', d['e2e']['value'], 'batched', d['batched_read']['value'], d['batched_read']['ms_per_launch'])
for c in ('cfg4','cfg5'):
    x=d[c]; print(c, x['single_frame']['value'], x['single_frame']['ms_per_launch'], x['batched_read']['value'])
"
exit 0
