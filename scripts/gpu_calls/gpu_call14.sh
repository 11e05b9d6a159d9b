mkdir -p gpurun_out/r2
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_inference_core.py -x -q -s > gpurun_out/r2/c14_pytest.txt 2>&1; echo "pytest rc=$?"; grep -E "amp vs|passed|failed|Error" gpurun_out/r2/c14_pytest.txt | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/c14_bench2.json 2> gpurun_out/r2/c14_bench2.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/r2/c14_bench2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2/c14_bench2.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
    s=d['sharded_cfg4']; print({k:s[k] for k in ('value','ms_per_step','efficiency_vs_same_run_single_gpu','speedup_vs_same_run_single_gpu','parity_ok','parity_rel_l2_max','per_rank_stage_us_max')}); print(s['single_gpu_same_run']); print(s['config']['exchange'])
except Exception as e:
    print('parse failed', e)
PY
EVAVOS_SHARD_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg4 > gpurun_out/r2/c14_cfg4_nccl.json 2> gpurun_out/r2/c14_cfg4_nccl.err; echo "nccl rc=$?"; tail -c 300 gpurun_out/r2/c14_cfg4_nccl.err
python - <<'PY'
import json
try:
    s=json.loads([l for l in open('gpurun_out/r2/c14_cfg4_nccl.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('NCCL engine', {k:s[k] for k in ('value','ms_per_step','efficiency_vs_same_run_single_gpu','parity_ok','per_rank_stage_us_max')})
except Exception as e:
    print('parse failed', e)
PY
