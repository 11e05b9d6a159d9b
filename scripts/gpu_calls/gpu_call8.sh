mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_memread.py tests/test_gpu_round2.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2/c8_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/c8_pytest.txt
tail -4 gpurun_out/r2/c8_pytest.txt
timeout 300 python scripts/stress_filter.py 12 > gpurun_out/r2/c8_stress.txt 2>&1; echo "stress rc=$?" >> gpurun_out/r2/c8_stress.txt
tail -2 gpurun_out/r2/c8_stress.txt
rm -f gpurun_out/r2/c8_filter.txt
for R in 1 2 3; do
  echo "== R=$R" >> gpurun_out/r2/c8_filter.txt
  FILTER_K=1 EVAVOS_SAMPLE_STRIDE=$R timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c8_filter.txt 2>&1
done
echo "== ld64 R=2" >> gpurun_out/r2/c8_filter.txt
FILTER_K=1 EVAVOS_SAMPLE_STRIDE=2 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_ld64.so timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c8_filter.txt 2>&1
cat gpurun_out/r2/c8_filter.txt
EVAVOS_SAMPLE_STRIDE=2 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_tr.so timeout 120 python scripts/trace_pass.py cfg4 > gpurun_out/r2/c8_trace_cfg4.txt 2>&1
grep -E "mean cycles|mma_ready ->" gpurun_out/r2/c8_trace_cfg4.txt
