#!/bin/bash
# call 40: compute-sanitizer over every kernel incl. the round-2 ones
mkdir -p gpurun_out/r2
python scripts/sanitize_small.py 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_small.py > gpurun_out/r2/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r2/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_small.py > gpurun_out/r2/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r2/sanitizer_racecheck.log
exit 0
