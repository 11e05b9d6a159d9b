#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/c32_bench_n8.json 2> gpurun_out/r2/c32_bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/c32_bench_n8.json') if l.startswith('{')][-1])
print('replicas', d['value'])
s=d['sharded_cfg4']; print('sharded', s['value'], s['efficiency_vs_same_run_single_gpu'], s['parity_ok'], s['per_rank_stage_us_max'], s['single_gpu_same_run'])
for k,v in d.get('hybrid_cfg4',{}).items(): print('hybrid', k, v['value'], v['efficiency_vs_same_run_single_gpu'], v['parity_ok'], v['per_rank_stage_us_max'])
PY
exit 0
