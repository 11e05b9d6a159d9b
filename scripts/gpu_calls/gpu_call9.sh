mkdir -p gpurun_out/r2
rm -f gpurun_out/r2/c9_filter.txt
for V in p2 p8 s10; do
  echo "== $V R=2" >> gpurun_out/r2/c9_filter.txt
  FILTER_K=1 EVAVOS_SAMPLE_STRIDE=2 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_$V.so timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c9_filter.txt 2>&1
done
cat gpurun_out/r2/c9_filter.txt
EVAVOS_SAMPLE_STRIDE=2 EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_trp8.so timeout 120 python scripts/trace_pass.py cfg4 > gpurun_out/r2/c9_trace_cfg4.txt 2>&1
grep -E "mean cycles|mma_ready ->" gpurun_out/r2/c9_trace_cfg4.txt
