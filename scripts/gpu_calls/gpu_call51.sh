#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2/c51_pytest.txt
timeout 300 python scripts/attention_time.py 2>&1 | tee gpurun_out/r2/c51_attention_time.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_partial_tc -c 1 -o gpurun_out/r2/full_attention_tc python scripts/attention_once.py > gpurun_out/r2/ncu_attention_tc.log 2>&1
tail -2 gpurun_out/r2/ncu_attention_tc.log
exit 0
