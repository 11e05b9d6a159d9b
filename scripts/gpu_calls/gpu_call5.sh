mkdir -p gpurun_out/r2
for W in cfg2 cfg4; do
EVAVOS_SAMPLE_STRIDE=2 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_tr.so timeout 120 python scripts/trace_pass.py $W > gpurun_out/r2/c5_trace_$W.txt 2>&1
grep -E "kernel marks|phase marks|mean cycles" gpurun_out/r2/c5_trace_$W.txt
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/c5_bench.json 2> gpurun_out/r2/c5_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2/c5_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2/c5_bench.json').read().strip().splitlines()[-1])
    print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()})
    print('stages', d['stages_us']); print('step roof', d['roofline']['step']); print('dominant', d['roofline']['kernel'], d['roofline']['frac'])
    print('batched', d['batched_read']['value'], d['batched_read']['stages_us'])
    print('gpu_baseline', d['gpu_baseline'])
    for c in ('cfg4','cfg5'):
        x=d[c]; print(c, {k:(x[k]['value'], x[k]['stages_us'], x[k]['roofline']['step']['frac']) for k in ('single_frame','batched_read')} if 'single_frame' in x else x)
        print(c, x.get('gpu_baseline'))
except Exception as e:
    print('parse failed', e)
PY
