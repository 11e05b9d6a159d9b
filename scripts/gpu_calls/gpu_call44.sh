#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2/c44_bench_n4.json 2> gpurun_out/r2/c44_bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/c44_bench_n4.json') if l.startswith('{')][-1])
print('replicas', d['value'], 'e2e', d['e2e']['value'])
s=d['sharded_cfg4']; print('sharded', s['value'], s['efficiency_vs_same_run_single_gpu'], s['parity_ok'], s['per_rank_stage_us_max'])
for k,v in d.get('hybrid_cfg4',{}).items(): print('hybrid', k, v['value'], v['efficiency_vs_same_run_single_gpu'], v['parity_ok'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
exit 0
