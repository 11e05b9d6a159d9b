#!/bin/bash
# call 27: overflow decision moved behind the query loads (finalize back to its old time?) + full tests + bench
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2/c27_tests.txt 2>&1
tail -4 gpurun_out/r2/c27_tests.txt
timeout 300 python scripts/stress_filter.py > gpurun_out/r2/c27_stress.txt 2>&1; tail -1 gpurun_out/r2/c27_stress.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/c27_bench.json 2> gpurun_out/r2/c27_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/c27_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['stages_us'], d['e2e']['value'])
"
python scripts/profile_cfg3.py amp graphs 2>&1 | head -12 | cut -c1-150
exit 0
