#!/bin/bash
# call 23: CUDA-graph replay of the conv passes, sticky overflow hint, readout with shared rows
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_inference_core.py tests/test_gpu_round2.py tests/test_networks.py -x -q -m gpu > gpurun_out/r2/c23_tests.txt 2>&1
tail -15 gpurun_out/r2/c23_tests.txt
python scripts/profile_cfg3.py amp > gpurun_out/r2/c23_cfg3_amp.txt 2>&1
python scripts/profile_cfg3.py amp graphs > gpurun_out/r2/c23_cfg3_amp_graphs.txt 2>&1
python scripts/profile_cfg3.py cl graphs > gpurun_out/r2/c23_cfg3_fp32_cl_graphs.txt 2>&1
python scripts/profile_cfg3.py graphs > gpurun_out/r2/c23_cfg3_fp32_graphs.txt 2>&1
grep -h "wall per video" gpurun_out/r2/c23_cfg3_*.txt
python scripts/readout_hot.py > gpurun_out/r2/c23_readout_hot.txt 2>&1
cat gpurun_out/r2/c23_readout_hot.txt | tail -10
for p in 0 3 0 3; do FILTER_PATH=$p python scripts/filter_time.py cfg2 2>&1 | tail -1; done > gpurun_out/r2/c23_filter_ab.txt
cat gpurun_out/r2/c23_filter_ab.txt
exit 0
