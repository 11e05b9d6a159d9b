#!/bin/bash
# call 21: exact tiled pass for overflowed queries (tests) + cfg3 with folded BatchNorm / channels_last
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_memread.py tests/test_gpu_inference_core.py -x -q -m gpu > gpurun_out/r2/c21_tests.txt 2>&1
tail -5 gpurun_out/r2/c21_tests.txt
python scripts/profile_cfg3.py > gpurun_out/r2/c21_cfg3_fp32.txt 2>&1
python scripts/profile_cfg3.py cl > gpurun_out/r2/c21_cfg3_fp32_cl.txt 2>&1
python scripts/profile_cfg3.py amp > gpurun_out/r2/c21_cfg3_amp.txt 2>&1
python scripts/profile_cfg3.py amp nofold > gpurun_out/r2/c21_cfg3_amp_nofold.txt 2>&1
grep -h "wall per video" gpurun_out/r2/c21_cfg3_*.txt
python scripts/filter_time.py cfg2 cfg4 > gpurun_out/r2/c21_filter.txt 2>&1
cat gpurun_out/r2/c21_filter.txt
exit 0
