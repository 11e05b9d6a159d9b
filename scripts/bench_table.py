"""Markdown tables for DESIGN.md from a bench.py JSON line.   python scripts/bench_table.py gpurun_out/r2/bench.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])


def rows(tag, stages, roof, value, ms):
    out = []
    for name, e in roof["kernels"].items():
        work = f"{e['algorithmic_flops'] / 1e9:.1f} GF" if "algorithmic_flops" in e else f"{e['algorithmic_bytes'] / 1e6:.1f} MB"
        out.append(f"| {tag} | `{name}` | {e['bound']} | {work} | {e['us_per_launch']:.1f} | {e['achieved']:.0f} {e['unit']} | **{e['frac']:.2f}** |")
    st = roof["step"]
    out.append(f"| {tag} | whole step ({value:.0f} qf/s) | {st['bound']} | T_roof {st['t_roof_us']:.0f} us | {st['us']:.1f} | | **{st['frac']:.2f}** |")
    return out


lines = ["| workload | kernel | bound | algorithmic work per launch | us (events) | achieved | fraction of measured peak |", "|---|---|---|---|---|---|---|"]
lines += rows("cfg2, 1 frame", d["stages_us"], d["roofline"], d["value"] / d["n_gpus"], d["ms_per_step"])
b = d["batched_read"]
lines += rows("cfg2, 5 frames per launch", b["stages_us"], b["roofline"], b["value"] / d["n_gpus"], b["ms_per_launch"])
for c in ("cfg4", "cfg5"):
    if c in d and "single_frame" in d[c]:
        for k, lab in (("single_frame", "1 frame"), ("batched_read", "5 frames per launch")):
            x = d[c][k]
            lines += rows(f"{c}, {lab}", x["stages_us"], x["roofline"], x["value"], x["ms_per_launch"])
print("\n".join(lines))
print()
print(f"cfg2: {d['value']:.0f} qf/s ({d['ms_per_step'] * 1e3:.1f} us per step); e2e {d['e2e']['value']:.0f} qf/s ({d['e2e'].get('mode')}); "
      f"sync-every-step {d.get('e2e_sync_every_step', {}).get('value', 0):.0f}; stateless {d['e2e_full_upload']['value']:.0f}; "
      f"GPU baseline {d['gpu_baseline']['value']:.0f} qf/s (x{d['gpu_baseline']['speedup']:.0f}); CPU port {d.get('cpu_baseline', {}).get('value', 0):.1f} qf/s on {d.get('cpu_baseline', {}).get('cores')} cores")
for c in ("cfg4", "cfg5"):
    if c in d and "gpu_baseline" in d[c]:
        print(c, "GPU baseline", d[c]["gpu_baseline"])
if "sharded_cfg4" in d:
    s = d["sharded_cfg4"]
    print("sharded", {k: s[k] for k in ("value", "ms_per_step", "efficiency_vs_same_run_single_gpu", "parity_ok", "per_rank_stage_us_max")}, s["single_gpu_same_run"])
