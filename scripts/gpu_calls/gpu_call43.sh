#!/bin/bash
# call 43: final bench line + launch list of the same build
mkdir -p gpurun_out/r2
K='regex:score_select|finalize_kernel|readout_|aggregate_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 40 --csv --log-file gpurun_out/r2/launches_bench.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r2/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/c43_bench.json 2> gpurun_out/r2/c43_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/c43_bench.json') if l.startswith('{')][-1])
print('cfg2', d['value'], d['ms_per_step'], d['stages_us'], d['e2e']['value'], d['clocks'])
"
exit 0
