mkdir -p gpurun_out/r2
rm -f gpurun_out/r2/c13_filter.txt
timeout 300 python scripts/stress_filter.py 12 > gpurun_out/r2/c13_stress.txt 2>&1; echo "stress rc=$?"; tail -1 gpurun_out/r2/c13_stress.txt
for V in base e16 e32 e48; do
  if [ $V = base ]; then L=evavos_b200/libevavos_sm100.so; else L=evavos_b200/libevavos_sm100_$V.so; fi
  echo "== $V R=2" >> gpurun_out/r2/c13_filter.txt
  FILTER_K=1 EVAVOS_SAMPLE_STRIDE=2 EVAVOS_LIB=$PWD/$L timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c13_filter.txt 2>&1
done
echo "== base R=1" >> gpurun_out/r2/c13_filter.txt
FILTER_K=1 EVAVOS_SAMPLE_STRIDE=1 timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c13_filter.txt 2>&1
cat gpurun_out/r2/c13_filter.txt
timeout 600 python -m pytest tests/test_gpu_memread.py tests/test_gpu_round2.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2/c13_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2/c13_pytest.txt
