set -x
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/r1_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"score_select|brute_select|finalize_kernel|readout_f32|aggregate_kernel" --launch-skip 15 -c 5 -f -o gpurun_out/r1_bench_kernels python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench2.log 2>&1; tail -2 gpurun_out/ncu_bench2.log | cut -c1-200
timeout 200 python bench.py --workload cfg5 2>/dev/null | tail -1 > gpurun_out/r1_bench_cfg5.json; cut -c1-300 gpurun_out/r1_bench_cfg5.json
timeout 200 python bench.py --workload cfg3 2>/dev/null | tail -1 > gpurun_out/r1_bench_cfg3.json; cut -c1-300 gpurun_out/r1_bench_cfg3.json
timeout 200 python bench.py --workload cfg4 2>/dev/null | tail -1 > gpurun_out/scale_cfg4_1.json; cut -c1-200 gpurun_out/scale_cfg4_1.json
