#!/bin/bash
# call 22: overflow hint + fused folded encoders + staged copies
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_memread.py tests/test_gpu_inference_core.py tests/test_networks.py -x -q -m gpu > gpurun_out/r2/c22_tests.txt 2>&1
tail -5 gpurun_out/r2/c22_tests.txt
python scripts/profile_cfg3.py > gpurun_out/r2/c22_cfg3_fp32.txt 2>&1
python scripts/profile_cfg3.py cl > gpurun_out/r2/c22_cfg3_fp32_cl.txt 2>&1
python scripts/profile_cfg3.py amp > gpurun_out/r2/c22_cfg3_amp.txt 2>&1
grep -h "wall per video" gpurun_out/r2/c22_cfg3_*.txt
python scripts/filter_time.py cfg2 cfg4 > gpurun_out/r2/c22_filter.txt 2>&1
cat gpurun_out/r2/c22_filter.txt
python -c "
import torch, evavos_b200 as ev
from evavos_b200.conv_opt import _probe_fused
d = torch.device('cuda:0')
for dt in (torch.float32, torch.bfloat16):
    for cl in (False, True):
        print('fused conv ops', dt, 'channels_last' if cl else 'nchw', _probe_fused(d, dt, cl))
"
exit 0
