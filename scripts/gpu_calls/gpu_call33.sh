#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2/c33_tests.txt 2>&1
tail -4 gpurun_out/r2/c33_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/c33_bench.json 2> gpurun_out/r2/c33_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
exit 0
