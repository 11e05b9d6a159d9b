#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_partial -o gpurun_out/r2/full_attention python scripts/attention_once.py > gpurun_out/r2/ncu_attention.log 2>&1
tail -3 gpurun_out/r2/ncu_attention.log
exit 0
