mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2/c1_smi.txt
for v in base e3 e7; do
  if [ $v = base ]; then L=evavos_b200/libevavos_sm100.so; else L=evavos_b200/libevavos_sm100_$v.so; fi
  echo "== $v" >> gpurun_out/r2/c1_filter.txt
  FILTER_K=1 EVAVOS_LIB=$PWD/$L timeout 200 python scripts/filter_time.py cfg2 cfg4 cfg5 >> gpurun_out/r2/c1_filter.txt 2>&1
done
for v in tr tr3 tr7; do
  echo "== $v" >> gpurun_out/r2/c1_trace.txt
  EVAVOS_LIB=$PWD/evavos_b200/libevavos_sm100_$v.so timeout 120 python scripts/trace_pass.py cfg4 >> gpurun_out/r2/c1_trace.txt 2>&1
done
cat gpurun_out/r2/c1_filter.txt
