mkdir -p gpurun_out/r2
rm -f gpurun_out/r2/c12_filter.txt
for R in 1 2; do
  echo "== base R=$R" >> gpurun_out/r2/c12_filter.txt
  FILTER_K=1 EVAVOS_SAMPLE_STRIDE=$R timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c12_filter.txt 2>&1
done
echo "== cl2 stress" >> gpurun_out/r2/c12_filter.txt
EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_cl2.so timeout 200 python scripts/stress_filter.py 12 >> gpurun_out/r2/c12_filter.txt 2>&1; echo "cl2 stress rc=$?" >> gpurun_out/r2/c12_filter.txt
for R in 1 2; do
  echo "== cl2 R=$R" >> gpurun_out/r2/c12_filter.txt
  FILTER_K=1 EVAVOS_SAMPLE_STRIDE=$R EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_cl2.so timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c12_filter.txt 2>&1
done
cat gpurun_out/r2/c12_filter.txt
EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_cl2.so timeout 600 python -m pytest tests/test_gpu_memread.py tests/test_gpu_round2.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2/c12_pytest_cl2.txt 2>&1; echo "cl2 pytest rc=$?"; tail -3 gpurun_out/r2/c12_pytest_cl2.txt
EVAVOS_SAMPLE_STRIDE=2 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_tr.so timeout 120 python scripts/trace_pass.py cfg2 > gpurun_out/r2/c12_trace_cfg2.txt 2>&1
grep -E "kernel marks|phase marks" gpurun_out/r2/c12_trace_cfg2.txt
