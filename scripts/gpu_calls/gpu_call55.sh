#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 > gpurun_out/r2/c55_bench_cfg3.json 2> gpurun_out/r2/c55_bench_cfg3.err; echo "cfg3 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2/c55_bench_cfg3.json') if l.startswith('{')][-1])
print('cfg3', d['value'], 'plain', d['plain_engine']['value'], 'amp', d['amp_variant']['value'], 'untail', d['without_fused_decoder_tails']['value'], 'ref', d['gpu_baseline'].get('value'))
for r in d['per_video_ms']: print(r)
PY
exit 0
