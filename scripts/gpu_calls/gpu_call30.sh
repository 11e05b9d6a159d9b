#!/bin/bash
mkdir -p gpurun_out/r2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/c30_bench_n2.json 2> gpurun_out/r2/c30_bench_n2.err; echo "bench n2 rc=$?"
grep -c "^{" gpurun_out/r2/c30_bench_n2.json
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
exit 0
