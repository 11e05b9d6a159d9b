mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/c2_pytest.txt
tail -15 gpurun_out/r2/c2_pytest.txt
timeout 300 python scripts/stress_filter.py 12 > gpurun_out/r2/c2_stress.txt 2>&1; echo "stress rc=$?" >> gpurun_out/r2/c2_stress.txt
tail -3 gpurun_out/r2/c2_stress.txt
for R in 1 2 3 4; do
  echo "== R=$R" >> gpurun_out/r2/c2_filter.txt
  FILTER_K=1 EVAVOS_SAMPLE_STRIDE=$R timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c2_filter.txt 2>&1
done
cat gpurun_out/r2/c2_filter.txt
echo "== e8 (no hit path) R=2" >> gpurun_out/r2/c2_filter.txt
FILTER_K=1 EVAVOS_SAMPLE_STRIDE=2 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_e8.so timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c2_filter.txt 2>&1
tail -4 gpurun_out/r2/c2_filter.txt
EVAVOS_SAMPLE_STRIDE=2 TRACE_I0=130 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_tr.so timeout 120 python scripts/trace_pass.py cfg4 > gpurun_out/r2/c2_trace_cfg4.txt 2>&1
