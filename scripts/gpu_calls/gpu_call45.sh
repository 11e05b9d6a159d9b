#!/bin/bash
mkdir -p gpurun_out/r2
nvidia-smi topo -m 2>/dev/null | head -12
nproc
for b in 1 0; do
EVAVOS_BIND=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2951$b bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2/c45_bench_n4_b$b.json 2> gpurun_out/r2/c45_b$b.err; echo "bind=$b rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2/c45_bench_n4_b$b.json') if l.startswith('{')][-1])
print('bind=$b', d['host_binding'], 'replicas', d['value'], 'e2e', d['e2e']['value'], d['e2e_sync_every_step']['value'])
PY
done
exit 0
