#!/bin/bash
# call 20b: where cfg3 spends its time
mkdir -p gpurun_out/r2
python scripts/profile_cfg3.py > gpurun_out/r2/c20_cfg3_fp32.txt 2>&1
python scripts/profile_cfg3.py amp > gpurun_out/r2/c20_cfg3_amp.txt 2>&1
python scripts/profile_cfg3.py amp bench > gpurun_out/r2/c20_cfg3_amp_bench.txt 2>&1
python scripts/profile_cfg3.py bench > gpurun_out/r2/c20_cfg3_bench.txt 2>&1
exit 0
