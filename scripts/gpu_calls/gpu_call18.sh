mkdir -p gpurun_out/r2
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2/bench_n$N.json 2> gpurun_out/r2/bench_n$N.err; echo "bench rc=$?"
grep -v "OMP_NUM\|\*\*\*\|^$" gpurun_out/r2/bench_n$N.err | tail -5 | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_n$N.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['mode'])
s=d['sharded_cfg4']; print({k:s[k] for k in ('value','ms_per_step','efficiency_vs_same_run_single_gpu','speedup_vs_same_run_single_gpu','parity_ok','parity_rel_l2_max','per_rank_stage_us_max')}); print(s['single_gpu_same_run'])
PY
EVAVOS_SHARD_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --workload cfg4 > gpurun_out/r2/cfg4_nccl_n$N.json 2> gpurun_out/r2/cfg4_nccl_n$N.err; echo "nccl rc=$?"
python - <<PY
import json
try:
    s=json.loads([l for l in open('gpurun_out/r2/cfg4_nccl_n$N.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('NCCL engine', {k:s[k] for k in ('value','ms_per_step','efficiency_vs_same_run_single_gpu','parity_ok','per_rank_stage_us_max')})
except Exception as e:
    print('parse failed', e)
PY
