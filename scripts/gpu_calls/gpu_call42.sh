#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/stress_filter.py 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/c42_bench.json 2> gpurun_out/r2/c42_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/c42_bench.json') if l.startswith('{')][-1])
print('cfg2', d['value'], d['ms_per_step'])
for c in ('cfg4','cfg5'):
    x=d[c]; print(c, x['single_frame']['value'], x['batched_read']['value'], x['batched_read']['stages_us'])
"
exit 0
