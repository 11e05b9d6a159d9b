#!/bin/bash
mkdir -p gpurun_out/r2
python scripts/profile_cfg3.py cl graphs 2>&1 | grep -v Warn | head -34 | cut -c1-230 | tee gpurun_out/r2/c54_profile_cfg3_default.txt
exit 0
