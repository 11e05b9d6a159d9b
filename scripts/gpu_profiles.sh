# Round-2 ncu evidence.  Keep gpurun_out/ small (<= 64 MiB are copied back): few launches per capture.
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_memread.py tests/test_gpu_round2.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2/prof_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2/prof_pytest.txt
K='regex:score_select|finalize_kernel|readout_|aggregate_kernel'
for W in cfg2 cfg4 cfg5; do
  # the LAST step of the workload: 4 reads x (filter, finalize, readout, aggregate) = 16 matching launches; skip 12
  SRC=""; if [ $W = cfg5 ]; then SRC="--import-source on"; fi
  timeout 400 ncu --set full $SRC --clock-control none -k "$K" -s 12 -c 4 -f -o gpurun_out/r2/full_$W python scripts/profile_step.py $W > gpurun_out/r2/ncu_$W.log 2>&1; echo "ncu $W rc=$?"
done
timeout 400 ncu --set full --clock-control none -k 'regex:write_keys|write_values' -s 8 -c 4 -f -o gpurun_out/r2/full_append python scripts/profile_step.py cfg2 > gpurun_out/r2/ncu_append.log 2>&1; echo "ncu append rc=$?"
timeout 400 ncu --set full --clock-control none -k 'regex:attention_|topk_merge|jf_|argmax' -c 12 -f -o gpurun_out/r2/full_extras python scripts/profile_step.py extras > gpurun_out/r2/ncu_extras.log 2>&1; echo "ncu extras rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 40 --csv --log-file gpurun_out/r2/launches_bench.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r2/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
du -sh gpurun_out/r2; ls -la gpurun_out/r2/*.ncu-rep | awk '{print $5, $9}'
