#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/r2/c31_tests.txt 2>&1; tail -5 gpurun_out/r2/c31_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/c31_bench_n2.json 2> gpurun_out/r2/c31_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2/c31_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/c31_bench_n2.json') if l.startswith('{')][-1])
s=d['sharded_cfg4']; print('sharded', s['value'], s['efficiency_vs_same_run_single_gpu'], s['parity_ok'], s['per_rank_stage_us_max'])
for k,v in d.get('hybrid_cfg4',{}).items(): print('hybrid', k, v['value'], v['efficiency_vs_same_run_single_gpu'], v['parity_ok'], v['per_rank_stage_us_max'])
PY
exit 0
