#!/bin/bash
# call 25: incremental loop state in the filter epilogue
mkdir -p gpurun_out/r2
timeout 300 python scripts/stress_filter.py > gpurun_out/r2/c25_stress.txt 2>&1; echo "stress rc=$?"
tail -2 gpurun_out/r2/c25_stress.txt
FILTER_K=1 timeout 600 python scripts/filter_time.py cfg2 cfg4 cfg5 > gpurun_out/r2/c25_filter.txt 2>&1
cat gpurun_out/r2/c25_filter.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2/c25_tests.txt 2>&1
tail -4 gpurun_out/r2/c25_tests.txt
exit 0
