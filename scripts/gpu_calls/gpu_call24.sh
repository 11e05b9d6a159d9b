#!/bin/bash
# call 24: 4 MMAs per tile + norm term added by the epilogue, A/B against the 5-MMA build of the same tree
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2/c24_tests.txt 2>&1
tail -6 gpurun_out/r2/c24_tests.txt
for lib in evavos_b200/libevavos_k80.so evavos_b200/libevavos_sm100.so evavos_b200/libevavos_k80.so evavos_b200/libevavos_sm100.so; do
  echo "== $lib"; EVAVOS_LIB=$lib FILTER_K=1 timeout 600 python scripts/filter_time.py cfg2 cfg4 cfg5 2>&1 | grep "^\["
done > gpurun_out/r2/c24_filter_ab.txt 2>&1
cat gpurun_out/r2/c24_filter_ab.txt
exit 0
