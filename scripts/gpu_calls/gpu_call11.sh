mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c11_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/c11_pytest.txt
tail -6 gpurun_out/r2/c11_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/c11_smoke.txt 2>&1; tail -2 gpurun_out/r2/c11_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/c11_bench.json 2> gpurun_out/r2/c11_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2/c11_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/c11_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'full', d['e2e_full_upload']['value'])
print('stages', d['stages_us']); print('step', d['roofline']['step']['frac'], 'dominant', d['roofline']['kernel'], d['roofline']['frac'])
print('batched', d['batched_read']['value'], d['batched_read']['stages_us'], d['batched_read']['roofline']['step']['frac'])
print('gpu_baseline', d['gpu_baseline']['value'], d['gpu_baseline']['speedup'], 'cpu', d['cpu_baseline']['value'])
for c in ('cfg4','cfg5'):
    x=d[c]
    if 'single_frame' in x:
        for k in ('single_frame','batched_read'):
            print(c,k,x[k]['value'], x[k]['stages_us'], 'step frac', x[k]['roofline']['step']['frac'], 'filter frac', x[k]['roofline']['kernels']['score_select_kernel']['frac'])
        print(c,'gpu_baseline', x.get('gpu_baseline'))
    else: print(c, x)
PY
