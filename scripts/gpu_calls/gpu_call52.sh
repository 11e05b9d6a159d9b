#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2/c52_pytest.txt
timeout 300 python scripts/attention_time.py 2>&1 | tee gpurun_out/r2/c52_attention_time.txt
exit 0
