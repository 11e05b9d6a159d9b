#!/bin/bash
mkdir -p gpurun_out/r2
for lib in evavos_b200/libevavos_sm100.so evavos_b200/libevavos_sm100_nosink.so evavos_b200/libevavos_sm100.so evavos_b200/libevavos_sm100_nosink.so; do
  echo "== $lib"; EVAVOS_LIB=$lib timeout 600 python scripts/filter_time.py cfg2 cfg4 2>&1 | grep "^\["
done > gpurun_out/r2/c28_sink_ab.txt 2>&1
cat gpurun_out/r2/c28_sink_ab.txt
exit 0
