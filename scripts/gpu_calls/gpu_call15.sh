mkdir -p gpurun_out/r2
K='regex:score_select|finalize_kernel|readout_|aggregate_kernel|write_keys|write_values|attention_|topk_merge|jf_|argmax'
for W in cfg2 cfg4 cfg5 extras; do
  timeout 900 ncu --set full --import-source on --clock-control none -k "$K" -f -o gpurun_out/r2/c15_full_$W python scripts/profile_step.py $W > gpurun_out/r2/c15_ncu_$W.log 2>&1; echo "ncu $W rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:score_select|finalize_kernel|readout_|aggregate_kernel' -c 40 --csv --log-file gpurun_out/r2/c15_launches_bench.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r2/c15_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/c15_bench.json 2> gpurun_out/r2/c15_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/c15_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'stages', d['stages_us'], 'batched', d['batched_read']['value'], d['batched_read']['stages_us'])
PY
for A in 0 1; do
EVAVOS_AMP=$A timeout 600 python bench.py --workload cfg3 --steps 6 --warmup 2 > gpurun_out/r2/c15_cfg3_amp$A.json 2> gpurun_out/r2/c15_cfg3_amp$A.err; echo "cfg3 amp=$A rc=$?"; cut -c1-200 gpurun_out/r2/c15_cfg3_amp$A.json
done
ls -la gpurun_out/r2/*.ncu-rep | awk '{print $5, $9}'
