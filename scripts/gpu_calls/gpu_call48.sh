#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_decoder_ops.py tests/test_gpu_inference_core.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2/c48_pytest.txt
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 > gpurun_out/r2/c48_bench_cfg3.json 2> gpurun_out/r2/c48_bench_cfg3.err; echo "cfg3 rc=$?"
tail -3 gpurun_out/r2/c48_bench_cfg3.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2/c48_bench_cfg3.json') if l.startswith('{')][-1])
print('cfg3', d['value'], 'plain', d['plain_engine']['value'], 'amp', d['amp_variant']['value'], 'untail', d['without_fused_decoder_tails']['value'])
print('gpu_baseline', d['gpu_baseline'])
PY
exit 0
