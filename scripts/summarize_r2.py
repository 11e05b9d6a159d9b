"""Round-2 ncu summaries.

    python scripts/summarize_r2.py full <rep> <workload> [<rep> <workload> ...]   -> profiles/r2_kernels_ncu.md, r2_traffic.json
    python scripts/summarize_r2.py launches <csv> <bench.json>                     -> profiles/r2_launches_bench.md (+ .csv copy)
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"), ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short_name(n):
    return n.split("(")[0].split("::")[-1]


def full(args):
    lines = ["# ncu summaries, round 2 (`--set full --clock-control none --import-source on`; per launch, cold cache, serialised)\n",
             "Captured with `scripts/profile_step.py <workload>` (banks built by frame appends, 4 fused reads + aggregations).",
             "Times under ncu are not bench values; the metrics explain them.\n"]
    traffic = {"source": "ncu --set full captures of scripts/profile_step.py, round 2: dram__bytes_read.sum + dram__bytes_write.sum per launch (last captured launch of each kernel)"}
    for rep, workload in zip(args[0::2], args[1::2]):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        last = {}
        for r in rows[2:]:
            last[short_name(r[hdr.index("Kernel Name")])] = r       # keep the last (warm) launch of every kernel
        lines.append(f"\n## {workload} (`{os.path.basename(rep)}`)\n")
        for name, r in last.items():
            lines.append(f"\n### {name}\n\n| metric | value |\n|---|---|")
            for key, label in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    lines.append(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
            try:
                b = sum(float(r[hdr.index(k)].replace(",", "")) * UNIT.get(units[hdr.index(k)], 1)
                        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                key = "readout_kernel" if name.startswith("readout") else ("aggregate_kernel" if name.startswith("aggregate") else name)
                traffic.setdefault(workload, {})[key] = b
            except Exception:
                pass
    open("profiles/r2_kernels_ncu.md", "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open("profiles/r2_traffic.json", "w"), indent=1)
    print("wrote profiles/r2_kernels_ncu.md, profiles/r2_traffic.json")


def launches(src, bench):
    shutil.copy(src, "profiles/r2_launches_bench.csv")
    text = open(src).read().splitlines()
    start = next(i for i, l in enumerate(text) if l.startswith('"ID"'))
    rows = []
    for r in csv.DictReader(io.StringIO("\n".join(text[start:]))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v, unit = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
            us = v / 1e3 if unit in ("nsecond", "ns") else v if unit in ("usecond", "us") else v * 1e3
            rows.append((short_name(r["Kernel Name"]), us))
    seq = ("score_select", "finalize", "readout", "aggregate")
    names = [n for n, _ in rows]
    idx = [i for i in range(len(rows) - 3) if all(names[i + j].startswith(s) for j, s in enumerate(seq))]
    tab = {}
    for i in idx[3:]:
        for n, us in rows[i:i + 4]:
            tab.setdefault(n, []).append(us)
    tot = sum(sum(v) / len(v) for v in tab.values())
    d = json.load(open(bench))
    out = ["# Launch list of `python bench.py` (first steps of its timed loop) under `ncu --metrics gpu__time_duration.sum --clock-control none`\n",
           "Full CSV: `profiles/r2_launches_bench.csv` (cold-cache, serialised: compare SHARES, not absolutes).\n",
           f"One step (query frame, cfg2), mean of {len(idx[3:])} steps after 3 warm-up steps:\n", "| kernel | us (ncu) | share |", "|---|---|---|"]
    for n, v in tab.items():
        m = sum(v) / len(v)
        out.append(f"| {n} | {m:.1f} | {100 * m / tot:.0f}% |")
    out.append(f"| total | {tot:.1f} | |")
    st = d.get("stages_us", {})
    s_tot = sum(st.values()) or 1.0
    out.append(f"\nFree-running bench (`profiles/r2_bench_line.json`): {d['ms_per_step'] * 1e3:.1f} us per step; its instrumented pass "
               "(CUDA events between the kernels) gives " + ", ".join(f"{k} {v:.1f} us ({100 * v / s_tot:.0f}%)" for k, v in st.items()) + ".")
    out.append("Four of this repo's kernels per step (`gpu_launches` = 4 x steps) plus one memset node (the filter's grid-barrier counters).")
    open("profiles/r2_launches_bench.md", "w").write("\n".join(out) + "\n")
    json.dump(d, open("profiles/r2_bench_line.json", "w"))
    print("\n".join(out))


if __name__ == "__main__":
    (full if sys.argv[1] == "full" else lambda a: launches(*a))(sys.argv[2:])
