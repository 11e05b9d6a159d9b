#!/bin/bash
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_decoder_ops.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2/c58_pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"upsample2x" -o gpurun_out/r2/full_decoder_ops2 python scripts/decoder_ops_once.py > gpurun_out/r2/ncu_decoder_ops2.log 2>&1
tail -2 gpurun_out/r2/ncu_decoder_ops2.log
exit 0
