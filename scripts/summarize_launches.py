"""profiles/r1_launches_bench.md from the ncu launch list of `bench.py --steps 2 --warmup 3` (see scripts/round_end_job.sh).

    python scripts/summarize_launches.py gpurun_out/r1_launches_bench.csv gpurun_out/r1_bench.json
"""
import csv
import io
import json
import shutil
import sys

src, bench = sys.argv[1], sys.argv[2]
shutil.copy(src, "profiles/r1_launches_bench.csv")
lines = open(src).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = []
for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v, unit = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        us = v / 1e3 if unit in ("nsecond", "ns") else v if unit in ("usecond", "us") else v * 1e3
        rows.append((r["Kernel Name"].split("(")[0].split("::")[-1], us))
names = [n for n, _ in rows]
seq = ("score_select", "brute", "finalize", "readout_f32", "aggregate_kernel")
idx = [i for i in range(len(rows) - 4) if all(names[i + j].startswith(s) for j, s in enumerate(seq))]
tab = {}
for i in idx[3:5]:   # warm-up steps 1-3, then the two timed steps of the device-resident loop
    for n, us in rows[i:i + 5]:
        tab.setdefault(n, []).append(us)
tot = sum(sum(v) / len(v) for v in tab.values())
d = json.load(open(bench))
ro = d["roofline"]
out = ["# Launch list of `python bench.py --steps 2 --warmup 3` under `ncu --metrics gpu__time_duration.sum --clock-control none`\n",
       "Full CSV: `profiles/r1_launches_bench.csv` (cold-cache, serialised, lower clocks than a free run: compare SHARES, not absolutes).\n",
       "One step (query frame) of the timed region (mean of its two steps):\n", "| kernel | us (ncu) | share |", "|---|---|---|"]
for n, v in tab.items():
    m = sum(v) / len(v)
    out.append(f"| {n} | {m:.1f} | {100 * m / tot:.0f}% |")
out.append(f"| total | {tot:.1f} | |")
out.append(f"\nFree-running bench (`profiles/r1_bench_line.json`): {d['ms_per_step'] * 1e3:.1f} us per step, readout "
           f"{ro['us_per_launch']:.1f} us = {100 * ro['share_of_step']:.0f}% of the step (CUDA events in the timed region), "
           "consistent with the ncu share above.")
out.append("Five of this repo's kernels per step (`gpu_launches` = 5 x steps) plus one 52-byte memset node (the filter's "
           "grid-barrier counters).")
open("profiles/r1_launches_bench.md", "w").write("\n".join(out) + "\n")
json.dump(d, open("profiles/r1_bench_line.json", "w"))
print("\n".join(out))
