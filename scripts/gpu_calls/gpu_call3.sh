mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_memread.py tests/test_gpu_round2.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2/c3_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/c3_pytest.txt
tail -25 gpurun_out/r2/c3_pytest.txt
timeout 300 python scripts/stress_filter.py 12 > gpurun_out/r2/c3_stress.txt 2>&1; echo "stress rc=$?" >> gpurun_out/r2/c3_stress.txt
tail -3 gpurun_out/r2/c3_stress.txt
rm -f gpurun_out/r2/c3_filter.txt
for R in 1 2 3 4; do
  echo "== R=$R" >> gpurun_out/r2/c3_filter.txt
  FILTER_K=1 EVAVOS_SAMPLE_STRIDE=$R timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c3_filter.txt 2>&1
done
cat gpurun_out/r2/c3_filter.txt
FILTER_K=1 EVAVOS_SAMPLE_STRIDE=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'finalize_kernel|score_select_kernel' -s 4 -c 4 -f -o gpurun_out/r2/c3_prof python scripts/filter_time.py cfg2 > gpurun_out/r2/c3_ncu.log 2>&1
tail -3 gpurun_out/r2/c3_ncu.log
FILTER_K=1 EVAVOS_SAMPLE_STRIDE=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'finalize_kernel|score_select_kernel' -s 4 -c 2 -f -o gpurun_out/r2/c3_prof_cfg5 python scripts/filter_time.py cfg5 > gpurun_out/r2/c3_ncu5.log 2>&1
tail -3 gpurun_out/r2/c3_ncu5.log
