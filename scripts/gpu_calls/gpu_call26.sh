#!/bin/bash
# call 26: round-2 final ncu evidence + bench lines
mkdir -p gpurun_out/r2
K='regex:score_select|finalize_kernel|readout_|aggregate_kernel'
for W in cfg2 cfg4 cfg5; do
  SRC=""; if [ $W = cfg5 ]; then SRC="--import-source on"; fi
  timeout 400 ncu --set full $SRC --clock-control none -k "$K" -s 12 -c 4 -f -o gpurun_out/r2/full_$W python scripts/profile_step.py $W > gpurun_out/r2/ncu_$W.log 2>&1; echo "ncu $W rc=$?"
done
timeout 400 ncu --set full --clock-control none -k 'regex:overflow_exact|finalize_kernel' -s 4 -c 2 -f -o gpurun_out/r2/full_dense python scripts/profile_step.py dense > gpurun_out/r2/ncu_dense.log 2>&1; echo "ncu dense rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 40 --csv --log-file gpurun_out/r2/launches_bench.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r2/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2/c26_smoke.txt 2>&1; tail -1 gpurun_out/r2/c26_smoke.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/c26_bench.json 2> gpurun_out/r2/c26_bench.err; echo "bench rc=$?"
python bench.py --workload cfg3 --steps 6 --warmup 3 > gpurun_out/r2/c26_cfg3.json 2> gpurun_out/r2/c26_cfg3.err; echo "cfg3 rc=$?"
cut -c1-400 gpurun_out/r2/c26_cfg3.json
du -sh gpurun_out/r2; ls -la gpurun_out/r2/full_*.ncu-rep | awk '{print $5, $9}'
exit 0
