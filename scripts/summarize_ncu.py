"""Summarise an .ncu-rep (raw page) into profiles/<name>.md: the metrics DESIGN.md / bench.py cite.

    python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_top_kernels.md [workload]
Also refreshes profiles/readout_traffic.json (dram bytes per launch of the readout kernel) when present.
"""
import csv
import io
import json
import os
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
workload = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
]
lines = [f"# ncu summary of `{os.path.basename(rep)}` (--set full --clock-control none; per launch, cold cache, serialised)\n"]
traffic = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    short = name.split("(")[0].split("::")[-1]
    lines.append(f"\n## {short}\n\n| metric | value |\n|---|---|")
    vals = {}
    for key, label in want:
        if key in hdr:
            i = hdr.index(key)
            vals[key] = (r[i], units[i])
            lines.append(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
    if "readout" in short and "dram__bytes_read.sum" in vals:
        def to_bytes(v, u):
            x = float(v.replace(",", ""))
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        traffic.setdefault(workload, []).append(to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"]))
with open(out, "w") as f:
    f.write("\n".join(lines) + "\n")
if traffic:
    path = os.path.join(os.path.dirname(out), "readout_traffic.json")
    cur = json.load(open(path)) if os.path.exists(path) else {}
    for k, v in traffic.items():
        cur[k] = sum(v) / len(v)
    json.dump(cur, open(path, "w"), indent=1)
print("wrote", out, traffic)
