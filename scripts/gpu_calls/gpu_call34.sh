#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
for v in "" _roq4 _roq2 _roq16 "" _roq4; do
  lib=evavos_b200/libevavos_sm100$v.so
  echo "== $lib"; EVAVOS_LIB=$lib timeout 600 python scripts/filter_time.py cfg2 cfg4 cfg5 2>&1 | grep "^\["
done > gpurun_out/r2/c34_readout_q.txt 2>&1
cat gpurun_out/r2/c34_readout_q.txt
exit 0
