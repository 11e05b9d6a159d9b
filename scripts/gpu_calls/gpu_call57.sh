#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bias_residual|upsample2x" -o gpurun_out/r2/full_decoder_ops python scripts/decoder_ops_once.py > gpurun_out/r2/ncu_decoder_ops.log 2>&1
tail -2 gpurun_out/r2/ncu_decoder_ops.log
exit 0
