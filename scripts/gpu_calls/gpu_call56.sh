#!/bin/bash
# 2 GPUs: the sharded tests that need two devices, then the N = 2 bench line as the driver launches it
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2/c56_pytest_sharded.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/c56_bench_n2.json 2> gpurun_out/r2/c56_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2/c56_bench_n2.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2/c56_bench_n2.json') if l.startswith('{')][-1])
print('replicas', d['value'], 'e2e', d['e2e']['value'])
s=d.get('sharded_cfg4',{}); print('sharded', {k:s.get(k) for k in ('value','parity_ok','efficiency','speedup','parallelism')})
print('hybrid', json.dumps(d.get('hybrid_cfg4'))[:600])
PY
exit 0
