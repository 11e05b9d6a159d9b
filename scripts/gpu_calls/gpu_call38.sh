#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2/c38_tests.txt 2>&1
tail -4 gpurun_out/r2/c38_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/c38_bench.json 2> gpurun_out/r2/c38_bench.err; echo "bench rc=$?"
python bench.py --workload cfg3 --steps 6 --warmup 3 > gpurun_out/r2/c38_cfg3.json 2> gpurun_out/r2/c38_cfg3.err; echo "cfg3 rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/c38_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['stages_us'], d['e2e']['value'])
c=json.loads(open('gpurun_out/r2/c38_cfg3.json').read()); print(c['value'], c['plain_engine']['value'], c['amp_variant']['value'])
"
exit 0
