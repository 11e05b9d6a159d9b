mkdir -p gpurun_out/r2
rm -f gpurun_out/r2/c16_filter.txt
for V in base seqld ss ss4; do
  if [ $V = base ]; then L=evavos_b200/libevavos_sm100.so; else L=evavos_b200/libevavos_sm100_$V.so; fi
  echo "== $V" >> gpurun_out/r2/c16_filter.txt
  EVAVOS_LIB=$PWD/$L timeout 120 python scripts/stress_filter.py 6 >> gpurun_out/r2/c16_filter.txt 2>&1; echo "stress $V rc=$?" >> gpurun_out/r2/c16_filter.txt
  FILTER_K=1 EVAVOS_LIB=$PWD/$L timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c16_filter.txt 2>&1
done
cat gpurun_out/r2/c16_filter.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/c16_bench.json 2> gpurun_out/r2/c16_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2/c16_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/c16_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'stages', d['stages_us'])
print('e2e', d['e2e']['value'], d['e2e']['mode'], 'pipelined', d['e2e_pipelined']['value'], 'sync', d['e2e_sync_every_step']['value'])
print('batched', d['batched_read']['value'], d['batched_read']['stages_us'])
PY
