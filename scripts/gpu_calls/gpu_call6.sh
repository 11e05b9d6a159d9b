mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_memread.py tests/test_gpu_round2.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2/c6_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/c6_pytest.txt
tail -5 gpurun_out/r2/c6_pytest.txt
timeout 300 python scripts/stress_filter.py 12 > gpurun_out/r2/c6_stress.txt 2>&1; echo "stress rc=$?" >> gpurun_out/r2/c6_stress.txt
tail -3 gpurun_out/r2/c6_stress.txt
rm -f gpurun_out/r2/c6_filter.txt
for R in 1 2 3 4; do
  echo "== R=$R" >> gpurun_out/r2/c6_filter.txt
  FILTER_K=1 EVAVOS_SAMPLE_STRIDE=$R timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c6_filter.txt 2>&1
done
cat gpurun_out/r2/c6_filter.txt
for W in cfg2 cfg4; do
EVAVOS_SAMPLE_STRIDE=2 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_tr.so timeout 120 python scripts/trace_pass.py $W > gpurun_out/r2/c6_trace_$W.txt 2>&1
grep -E "kernel marks|phase marks" gpurun_out/r2/c6_trace_$W.txt
done
