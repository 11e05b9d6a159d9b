"""bench.py - memory-read query-frames/s of the B200 space-time memory read (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3|cfg4|cfg5]

A step is one query frame of the hot path on synthetic inputs: key affinity + top-k softmax + value readout for all
objects (one evavos_memread call: filter, finalize, readout kernels) followed by the soft aggregation across objects.
Default workload (BASELINE.json configs[1]): DAVIS-17 480p, 30x54 feature map, 3 objects, 20-frame bank.

ONE JSON line:
 * value      whole-job query-frames/s with the banks resident in HBM (CUDA events, max over ranks); four banks
              (together > L2) are rotated so no step finds its inputs in L2.  `batched_read` is the same workload the
              way the product issues it (inference_core.py: the mem_freq = 5 query frames between two memory appends
              share one launch).
 * roofline   one entry per kernel of the step (`kernels`), the whole step against SURVEY 8(d)'s T_roof (`step`), and
              at top level the DOMINANT (longest) kernel.  Kernel times come from a second, instrumented pass over the
              same steps (CUDA events between the kernels, recorded by the library: evavos_stage_timing); the event
              records defeat the kernels' programmatic-dependent-launch overlap, so they add up to more than
              ms_per_step.  Readout bytes count the DISTINCT value rows the step selected (measured).
 * gpu_baseline  the reference's own torch op sequence (oracle/torch_port.py: SGEMM + 3 elementwise passes, topk,
              scatter, one bmm per object, 6-8 elementwise aggregate launches) on CUDA tensors on the same B200 - the
              "existing GPU path" (SURVEY 8d).  Checker code, timed beside the product, never called by it.
 * e2e        the same step through the public API with pinned HOST buffers (streaming: the bank is engine state);
              `e2e_full_upload` is the stateless evavos_memread_host form.
 * cpu_baseline / --impl reference   oracle/torch_port.py on all host cores (/root/reference is not on the GPU box).
 * N == 1 also carries `cfg4` and `cfg5`: BASELINE.json configs[3] (unsharded, one GPU) and configs[4] (bf16 values)
   with their own stage times, rooflines and GPU baselines; and `attention_read`: the fusion path's attention read
    at the 480p and 1080p maps, both forms of csrc/attention.cu, beside the reference's torch ops.
 * --workload cfg3: `InferenceCore.interact` end to end (frames/s of the default engine, the plain engine, the bf16
   variant, the engine without the fused decoder tails, per-video median / min / max), and `gpu_baseline`: the
   reference's stock propagation loop (oracle/stock_engine.py) on the same GPU with the mask agreement between the two.
 * N > 1: ranks run independent videos (no collective, weak scaling: the headline line), and the line also carries
   `sharded_cfg4`: ONE 200-frame bank sharded along the memory axis over the N ranks (device-initiated exchange over
   NVLink peer memory; EVAVOS_SHARD_EXCHANGE=nccl for the library baseline) with per-rank stage times, the same run's
   single-GPU time of the same read, the efficiency between the two, and `parity_ok` (every rank's result against its
   own single-bank read); and `hybrid_cfg4`: the same read with the queries split as well (N / M query groups x M
   memory shards, M = N/2, N/4, 1 - evavos_b200.sharded.HybridShardedBank).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (CK, CV, T, H, W, K objects, seed, description)
    "cfg1": (64, 512, 5, 30, 54, 1, 1235, "DAVIS-17 480p (30x54), 1 object, 5-frame bank"),
    "cfg2": (64, 512, 20, 30, 54, 3, 1236, "DAVIS-17 480p (30x54), 3 objects, 20-frame bank, top-50 readout + soft aggregation"),
    "cfg4": (64, 512, 200, 30, 54, 1, 1238, "MOSE-style long video 480p, 1 object, 200-frame bank"),
    "cfg3": (64, 512, 32, 30, 54, 1, 1237, "independent synthetic 480p videos (32 frames, 1 object), full key/value encode + memory read + decode"),
    "cfg5": (64, 512, 50, 68, 120, 5, 1239, "1080p-equivalent feature map (68x120), 5 objects, 50-frame bank, bf16 value bank"),
}
TOP_K = 50
N_BANKS = 4
MEM_FREQ = 5              # query frames between two memory appends (inference_core.py:174): read in one launch
KERNELS_PER_STEP = 4      # score_select_kernel, finalize_kernel, readout kernel, aggregate_kernel (+ one memset node)


def synth(seed, ck, cv, t, h, w, k):
    g = torch.Generator().manual_seed(seed)
    mk = torch.randn(1, ck, t, h, w, generator=g)
    qk = torch.randn(1, ck, h, w, generator=g)
    mv = torch.randn(k, cv, t, h, w, generator=g)
    return mk, qk, mv


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "tc": float(p["bf16_tflops_sustained"]), "tc_burst": float(p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json; bf16 = sustained figure, the kernels run inside a long step)"}
    except Exception:
        return {"hbm": 6650.0, "tc": 1400.0, "tc_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_rate(cfg, steps, warmup, budget_s=150.0):
    """Reference dense torch path (oracle/torch_port.py) on all host cores; returns (qf/s, info)."""
    from oracle import torch_port as port
    ck, cv, t, h, w, k, seed, _ = cfg
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mk, qk, mv = synth(seed, ck, cv, t, h, w, k)
    prob = torch.rand(k, 1, h * 16, w * 16, generator=torch.Generator().manual_seed(4321))
    frac = 1.0
    hw = h * w

    def step(fr):
        cols = max(1, int(round(hw * fr)))
        q = qk.flatten(2)[:, :, :cols].reshape(1, ck, 1, cols)     # a slice of the query columns
        out = port.memory_read(mk, q, mv, TOP_K)
        agg = port.aggregate_wbg(prob, keep_bg=True)
        return out, agg

    t0 = time.perf_counter()
    step(1.0)
    t_one = time.perf_counter() - t0
    while frac > 1 / 64 and t_one * frac * (steps + warmup) > budget_s:
        frac /= 2
    for _ in range(max(0, warmup - 1)):
        step(frac)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(frac)
    dt = time.perf_counter() - t0
    rate = steps * frac / dt
    sample = (f"{steps} steps x {frac:g} of the {hw} query columns of one query frame (dense affinity "
              f"{t * hw}x{int(round(hw * frac))} fp32, topk, scatter, {k} bmm) + aggregate_wbg {k}x{h * 16}x{w * 16}")
    return rate, {"cores": cores, "kind": "port", "sample": sample, "ms_per_step": 1e3 * dt / steps, "frac": frac}


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    rate, info = cpu_reference_rate(cfg, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "memory-read query-frames/sec", "value": rate, "unit": "query-frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["ms_per_step"] / info["frac"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + cfg[7], "top_k": TOP_K},
        "cpu_baseline": {"value": rate, "unit": "query-frames/s", "cores": info["cores"], "kind": "port",
                         "sample": info["sample"]},
        "e2e": {"value": rate, "unit": "query-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- helpers (GPU)
def time_loop(fn, steps, warmup, stream, dev):
    """ms for `steps` calls of fn(i) after `warmup`, CUDA events on the launching stream."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        fn(i)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1)


def stage_pass(read_fn, agg_fn, steps, stream, dev):
    """Instrumented pass: mean us of {filter, finalize, readout} (library events) and of the aggregate kernel."""
    from evavos_b200 import _lib
    lib = _lib.load()
    steps = min(steps, 200)
    lib.evavos_stage_timing(1)
    ev = []
    for i in range(steps):
        read_fn(i)
        if agg_fn is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            agg_fn(i)
            b.record(stream)
            ev.append((a, b))
    ms = (ctypes.c_float * 4)()
    _lib.check(lib.evavos_stage_timing_read(ms))
    lib.evavos_stage_timing(0)
    torch.cuda.synchronize(dev)
    out = {"filter": ms[0] * 1e3, "finalize": ms[2] * 1e3, "readout": ms[3] * 1e3}
    if ev:
        out["aggregate"] = 1e3 * sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    return out


def rooflines(shape, stages_us, step_us, distinct_rows, val_bytes, pk, frames=1):
    """Per-kernel roofline entries + the whole step.  Returns (dominant, kernels, step)."""
    ck, cv, n_pos, hw, k_obj, agg_px = shape
    nq = hw * frames
    kernels = {}
    f_filter = 2.0 * n_pos * nq * ck
    us = stages_us["filter"]
    kernels["score_select_kernel"] = {
        "bound": "tensor", "us_per_launch": us, "algorithmic_flops": f_filter, "achieved": f_filter / us / 1e6,
        "peak": pk["tc"], "unit": "TFLOP/s", "frac": f_filter / us / 1e6 / pk["tc"],
        "note": "2*N*HW*CK: ONE pass of the 64-channel contraction; the kernel contracts 1 + 1/sample_stride sweeps and a 5th K=16 slice for -|k|^2/2"}
    b_fin = float(nq) * TOP_K * (4 * ck + 8 + 8)
    us = stages_us["finalize"]
    kernels["finalize_kernel"] = {
        "bound": "hbm", "us_per_launch": us, "algorithmic_bytes": b_fin, "achieved": b_fin / us / 1e3, "peak": pk["hbm"],
        "unit": "GB/s", "frac": b_fin / us / 1e3 / pk["hbm"],
        "note": "k exact key rows + k list entries + k outputs per query; latency-bound (dependent gathers), not bandwidth-bound"}
    b_ro = float(val_bytes) * k_obj * cv * distinct_rows + 4.0 * k_obj * cv * nq + 8.0 * TOP_K * nq
    us = stages_us["readout"]
    kernels["readout_kernel"] = {
        "bound": "hbm", "us_per_launch": us, "algorithmic_bytes": b_ro, "achieved": b_ro / us / 1e3, "peak": pk["hbm"],
        "unit": "GB/s", "frac": b_ro / us / 1e3 / pk["hbm"], "distinct_value_rows": int(distinct_rows),
        "note": "every DISTINCT selected value row once (measured count) + output once + (idx, weight) once"}
    if "aggregate" in stages_us:
        b_ag = float(2 * k_obj + 1) * agg_px * 4 * frames
        us = stages_us["aggregate"]
        kernels["aggregate_kernel"] = {
            "bound": "hbm", "us_per_launch": us / frames, "algorithmic_bytes": b_ag / frames, "achieved": b_ag / us / 1e3,
            "peak": pk["hbm"], "unit": "GB/s", "frac": b_ag / us / 1e3 / pk["hbm"]}
    # whole step, SURVEY 8(d): F_alg = key contraction + sparse readout; B_alg = each needed input once + outputs once
    f_alg = f_filter + 2.0 * k_obj * cv * TOP_K * nq
    s_key = 2.0 if val_bytes == 2 else 4.0
    b_alg = s_key * (n_pos * ck + nq * ck) + val_bytes * k_obj * cv * min(n_pos, TOP_K * nq) + 4.0 * k_obj * cv * nq
    b_alg_distinct = b_alg - val_bytes * k_obj * cv * (min(n_pos, TOP_K * nq) - distinct_rows)
    if "aggregate" in stages_us:
        b_alg += (2 * k_obj + 1) * agg_px * 4.0 * frames
        b_alg_distinct += (2 * k_obj + 1) * agg_px * 4.0 * frames
    t_tc, t_hbm = f_alg / pk["tc"] / 1e6, b_alg / pk["hbm"] / 1e3
    t_roof = max(t_tc, t_hbm)
    step = {"t_roof_us": t_roof, "t_tensor_us": t_tc, "t_hbm_us": t_hbm, "bound": "tensor" if t_tc > t_hbm else "hbm",
            "us": step_us, "frac": t_roof / step_us, "algorithmic_flops": f_alg, "algorithmic_bytes": b_alg,
            "frac_with_distinct_rows": max(t_tc, b_alg_distinct / pk["hbm"] / 1e3) / step_us,
            "note": "SURVEY.md 8(d): max(F_alg / P_tc, B_alg / BW_hbm) over the measured (un-instrumented) step time"}
    name = max(kernels, key=lambda n: kernels[n]["us_per_launch"] * (frames if n == "aggregate_kernel" else 1))
    dom = dict(kernels[name])
    dom["kernel"] = name
    return dom, kernels, step


def build_banks(cfg, dev, n_banks, seed_shift=0, bf16=False, on_device=False):
    import evavos_b200 as ev
    ck, cv, t, h, w, k, seed, _ = cfg
    banks, queries = [], []
    for b in range(n_banks):
        vd = torch.bfloat16 if bf16 else torch.float32
        bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, value_dtype=vd, keep_reference_layout=False)
        if on_device:     # large banks: synthesised frame by frame on the device
            g = torch.Generator(device=dev).manual_seed(seed + 100 * b + seed_shift)
            for f in range(t):
                kf = torch.randn(1, ck, h, w, generator=g, device=dev)
                vf = torch.randn(k, cv, 1, h, w, generator=g, device=dev)
                if bf16:
                    kf, vf = kf.to(torch.bfloat16).float(), vf.to(torch.bfloat16).float()
                bank.append(kf, vf)
            q = torch.randn(1, ck, MEM_FREQ, h, w, generator=g, device=dev)
            q = q.to(torch.bfloat16).float() if bf16 else q
        else:
            mk, qk, mv = synth(seed + 100 * b + seed_shift, ck, cv, t, h, w, k)
            for f in range(t):  # built the way do_pass builds it: one append per memory frame
                bank.append(mk[:, :, f].to(dev), mv[:, :, f:f + 1].to(dev))
            g = torch.Generator().manual_seed(seed + 100 * b + seed_shift + 7)
            q = torch.cat([qk.unsqueeze(2), torch.randn(1, ck, MEM_FREQ - 1, h, w, generator=g)], 2).to(dev)
        banks.append(bank)
        queries.append(q)
    return banks, queries


def gpu_baseline(cfg, bank_tensors, steps, dev, stream, with_aggregate=True):
    """The reference's own torch sequence on CUDA tensors (oracle/torch_port.py, bit-equal to the reference on CPU)."""
    from oracle import torch_port as port
    ck, cv, t, h, w, k, seed, _ = cfg
    mk, qk, mv = bank_tensors
    prob = torch.rand(k, 1, h * 16, w * 16, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = False      # the reference's default: true fp32 matmul / bmm

    def step(i):
        out = port.memory_read(mk, qk, mv, TOP_K)
        if with_aggregate:
            port.aggregate_wbg(prob, keep_bg=True)
        return out

    ms = time_loop(step, steps, 2, stream, dev)
    return {"value": steps / (ms * 1e-3), "unit": "query-frames/s", "ms_per_step": ms / steps, "steps": steps,
            "impl": "oracle/torch_port.py (the reference's ATen op sequence) on CUDA tensors, fp32, allow_tf32=False",
            "launches_per_step": "~%d" % (12 + k + (8 if with_aggregate else 0))}


def side_workload(name, args, dev, stream, pk, want_baseline=True):
    """cfg4 / cfg5 on one GPU: value, stage times, rooflines, GPU baseline (extra keys of the N=1 line)."""
    import evavos_b200 as ev
    cfg = WORKLOADS[name]
    ck, cv, t, h, w, k, seed, desc = cfg
    bf16 = name == "cfg5"
    hw, n_pos = h * w, t * h * w
    banks, queries = build_banks(cfg, dev, 2, bf16=bf16, on_device=True)
    steps = max(5, min(args.steps, 20))
    res = {"workload": name + ": " + desc, "dtype": "bf16 values + bf16-representable keys, fp32 accumulation" if bf16 else "f32",
           "memory_positions": n_pos, "queries_per_frame": hw, "objects": k,
           "l2": f"2 rotating banks of {banks[0].val_pm.numel() * banks[0].val_pm.element_size() / 1e9:.2f} GB values"}
    for frames, key in ((1, "single_frame"), (MEM_FREQ, "batched_read")):
        qs = [q[:, :, 0] if frames == 1 else q for q in queries]
        outs = [torch.empty((k, cv, frames * hw), dtype=torch.float32, device=dev) for _ in range(2)]

        def read(i):
            return ev.memory_read(banks[i % 2], qs[i % 2], TOP_K, out=outs[i % 2].view(k, cv, *qs[i % 2].shape[2:]))

        ms = time_loop(read, steps, 3, stream, dev)
        stages = stage_pass(read, None, steps, stream, dev)
        _, aff = ev.memory_read(banks[0], qs[0], TOP_K, want_readout=False, want_topk=True)
        distinct = int(torch.unique(aff.idx).numel())
        dom, kernels, step = rooflines((ck, cv, n_pos, hw, k, 0), stages, 1e3 * ms / steps, distinct, 2 if bf16 else 4, pk, frames)
        res[key] = {"value": frames * steps / (ms * 1e-3), "unit": "query-frames/s", "ms_per_launch": ms / steps,
                    "query_frames_per_launch": frames, "steps": steps, "stages_us": stages,
                    "roofline": {"dominant": dom["kernel"], "kernels": kernels, "step": step}}
    if want_baseline:
        # the same bank as dense reference-layout tensors (rebuilt from the shadow: position-major -> (C,T,H,W))
        b0 = banks[0]
        mk = b0.key_pm[:n_pos].t().reshape(1, ck, t, h, w).contiguous()
        mv = b0.val_pm[:, :n_pos].float().permute(0, 2, 1).reshape(k, cv, t, h, w).contiguous()
        try:
            res["gpu_baseline"] = gpu_baseline(cfg, (mk, queries[0][:, :, 0].contiguous(), mv), 3, dev, stream, with_aggregate=False)
            res["gpu_baseline"]["speedup_single_frame"] = res["single_frame"]["value"] / res["gpu_baseline"]["value"]
        except Exception as e:   # e.g. out of memory for the dense 13.3 GB affinity on a busy device
            res["gpu_baseline"] = {"unavailable": repr(e)[:200]}
        del mk, mv
    del banks, queries
    torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------------------------------------- sharded cfg4
def attention_workload(dev, stream, pk, steps=20):
    """The fusion path's attention read (AttentionMemory + get_attention's two vector-matrix products,
    prop_net.py:117-138, 198-211) at the 480p and the 1080p map, both forms of csrc/attention.cu, against the same ops
    of the reference in torch on this GPU (oracle.torch_port.attention_lowres: dense HW x HW softmax)."""
    import evavos_b200 as ev
    from oracle import torch_port
    F = torch.nn.functional
    res = {}
    for name, (h, w) in (("480p_30x54", (30, 54)), ("1080p_68x120", (68, 120))):
        g = torch.Generator().manual_seed(h)
        mk, qk = torch.randn(1, 64, 1, h, w, generator=g).to(dev), torch.randn(1, 64, h, w, generator=g).to(dev)
        pos = (torch.rand(1, 1, 16 * h, 16 * w, generator=g) > 0.5).float().to(dev)
        neg = 1 - pos
        vec = torch.cat([F.interpolate(pos, size=(h, w), mode="area").view(1, -1),
                         F.interpolate(neg, size=(h, w), mode="area").view(1, -1),
                         torch.rand(2, h * w, generator=g).to(dev)], 0)          # 4 mask rows: K = 1 (pos / neg per channel)
        row, old = {"scores": (h * w) ** 2, "mask_rows": 4}, os.environ.get("EVAVOS_ATTENTION_PATH")
        try:
            for form in ("simt", "tensor"):
                os.environ["EVAVOS_ATTENTION_PATH"] = form
                row[form + "_us"] = 1e3 * time_loop(lambda i: ev.attention_readout(mk, qk, vec), steps, 3, stream, dev) / steps
                row[form + "_out"] = ev.attention_readout(mk, qk, vec)
        finally:
            os.environ.pop("EVAVOS_ATTENTION_PATH", None) if old is None else os.environ.__setitem__("EVAVOS_ATTENTION_PATH", old)
        row["default_form"] = "tensor" if (h * w) ** 2 >= (1 << 22) else "simt"
        ref_fn = lambda i: torch_port.attention_lowres(mk, pos, neg, qk)
        row["torch_us"] = 1e3 * time_loop(ref_fn, 5, 2, stream, dev) / 5
        want = ref_fn(0).reshape(2, -1)
        row["max_abs_vs_torch"] = max(float((row.pop(f + "_out")[:2] - want).abs().max()) for f in ("simt", "tensor"))
        us = row[row["default_form"] + "_us"]
        row["speedup_vs_torch"] = row["torch_us"] / us
        row["useful_tflops"] = 2.0 * (h * w) ** 2 * 64 / us / 1e6
        row["frac_of_bf16_peak"] = row["useful_tflops"] / pk["tc"]
        res[name] = row
    return res


def run_sharded(args, cfg, rank, world, local_rank, dist=None, emit=True, memory_shards=None):
    """cfg4: one long bank sharded by frame over the ranks; strong scaling (total work fixed).

    memory_shards = M < world: the ranks form world / M query groups, each holding the whole bank sharded over its M
    ranks and answering its own slice of the queries (evavos_b200.sharded.HybridShardedBank); M = world (default) is
    the plain memory-axis sharded read.
    Returns the result dict (rank 0) - printed as its own line only with --workload cfg4."""
    import evavos_b200 as ev
    from evavos_b200.sharded import HybridShardedBank
    ck, cv, t, h, w, k, seed, desc = cfg
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    own_pg = False
    if world > 1 and dist is None:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
        own_pg = True
    hw, n_pos = h * w, t * h * w
    n_banks = 2
    steps = max(5, min(args.steps, 50))
    m_shards = world if memory_shards is None else int(memory_shards)
    banks, full, queries = [], [], []
    for b in range(n_banks):
        g = torch.Generator(device=dev).manual_seed(seed + 100 * b)      # the same stream of frames on every rank
        bank = HybridShardedBank(k, ck, cv, h, w, t, dev, m_shards)
        whole = ev.MemoryBank(k, ck, cv, h, w, t, dev, keep_reference_layout=False) if b == 0 else None
        for f in range(t):
            kf = torch.randn(1, ck, h, w, generator=g, device=dev)
            vf = torch.randn(k, cv, 1, h, w, generator=g, device=dev)
            bank.append(kf, vf)                                           # only the owner keeps it
            if whole is not None:
                whole.append(kf, vf)
        banks.append(bank)
        full.append(whole)
        queries.append(torch.randn(1, ck, MEM_FREQ, h, w, generator=g, device=dev))
    stream = torch.cuda.current_stream(dev)

    def step(i):
        # mem_freq = 5 query frames share one bank state (inference_core.py:174): one exchange for all of them;
        # every rank ends with the readout of the query slice it owns (the decoder consumes it where it is)
        return banks[i % n_banks].read(queries[i % n_banks], TOP_K)

    def step_single(i):
        return ev.memory_read(full[0], queries[i % n_banks], TOP_K)[0]

    # parity: every rank's sharded result against its own single-bank read of the same bank
    out_s, idx_s, w_s = banks[0].read(queries[0], TOP_K, return_topk=True)
    out_1, aff_1 = ev.memory_read(full[0], queries[0], TOP_K, want_topk=True)
    torch.cuda.synchronize(dev)
    qa, qb = banks[0].query_range(MEM_FREQ * hw)        # the queries this rank's group answers
    q0, q1 = banks[0].owned_slice(MEM_FREQ * hw)        # ... and the ones whose readout this rank ends up with
    ref_slice = out_1.reshape(k, cv, MEM_FREQ * hw)[:, :, q0:q1]
    rel = float(((out_s - ref_slice).norm() / ref_slice.norm()).item())
    same_idx = bool(torch.equal(idx_s, aff_1.idx[qa:qb]))
    parity = torch.tensor([1.0 if (rel < 1e-5 and same_idx) else 0.0, rel], device=dev, dtype=torch.float64)

    for i in range(max(3, args.warmup)):
        step(i)
    torch.cuda.synchronize(dev)
    if dist:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        step(i)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if dist:
        dist.barrier()
    elapsed_ms = e0.elapsed_time(e1)
    single_ms = time_loop(step_single, steps, 3, stream, dev)
    stage_us = banks[0].profile(queries[0], TOP_K, reps=10) if hasattr(banks[0], "profile") else {}
    if dist:
        tm = torch.tensor([elapsed_ms, single_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        elapsed_ms, single_ms = float(tm[0].item()), float(tm[1].item())
        dist.all_reduce(parity, op=dist.ReduceOp.MAX)   # worst rel error over the ranks
        ok = torch.tensor([1.0 if (rel < 1e-5 and same_idx) else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity_ok = bool(ok.item() > 0.5)
        st = torch.tensor([stage_us.get(n, 0.0) for n in sorted(stage_us)], device=dev, dtype=torch.float64)
        if st.numel():
            dist.all_reduce(st, op=dist.ReduceOp.MAX)
            stage_us = {n: float(v) for n, v in zip(sorted(stage_us), st.tolist())}
    else:
        parity_ok = bool(rel < 1e-5 and same_idx)
    res = None
    if rank == 0:
        value = MEM_FREQ * steps / (elapsed_ms * 1e-3)
        single = MEM_FREQ * steps / (single_ms * 1e-3)
        flops = 2.0 * n_pos * hw * ck * MEM_FREQ
        pk = peaks()
        res = {
            "metric": "memory-read query-frames/sec", "value": value, "unit": "query-frames/s", "n_gpus": world,
            "steps": steps, "ms_per_step": elapsed_ms / steps, "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg4: " + desc + (f", memory axis sharded by frame over {world} GPU(s)" if m_shards == world else
                                                      f", {world // m_shards} query groups x {m_shards} memory shards"), "top_k": TOP_K,
                       "memory_positions": n_pos, "queries_per_frame": hw, "objects": k, "query_frames_per_step": MEM_FREQ,
                       "parallelism": f"memory-axis shards x{world}" if m_shards == world else
                                      f"query groups x{world // m_shards}, memory-axis shards x{m_shards}",
                       "memory_shards": m_shards, "bank_fraction_per_gpu": 1.0 / m_shards,
                       "exchange": getattr(banks[0], "exchange_desc", "all-gather(top-k) + all-reduce(readout)")},
            "single_gpu_same_run": {"value": single, "ms_per_step": single_ms / steps,
                                    "note": "the same 5-frame read against the whole bank on ONE GPU, timed in this run"},
            "efficiency_vs_same_run_single_gpu": value / (world * single),
            "speedup_vs_same_run_single_gpu": value / single,
            "per_rank_stage_us_max": stage_us,
            "parity_ok": parity_ok, "parity_rel_l2_max": float(parity[1].item()),
            "roofline": {"bound": "tensor", "kernel": "whole sharded read", "achieved": flops / (elapsed_ms / steps * 1e-3) / 1e12 / world,
                         "peak": pk["tc"], "unit": "TFLOP/s", "frac": flops / pk["tc"] / 1e12 / world / (elapsed_ms / steps * 1e-3),
                         "note": "algorithmic flops 2*N*HW*CK per query frame over the measured step time, per GPU"},
        }
        if emit:
            line = dict(res)
            line.update({"warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
                         "gpu_launches": (3 + 2) * steps})
            print(json.dumps(line), flush=True)
    for b in banks:
        if hasattr(b, "close"):
            b.close()
    del banks, full, queries
    torch.cuda.empty_cache()
    if own_pg:
        dist.destroy_process_group()
    return res


# ----------------------------------------------------------------------------------------------- cfg3 (end to end)
def run_cfg3(args, cfg, rank, world, local_rank):
    """BASELINE.json configs[2]: end-to-end InferenceCore.interact on independent synthetic 480p videos, one process
    per GPU, no collective.  A step is one video (mask on frame 0, propagated to the other 31 frames); the conv
    encoders/decoder (PyTorch/cuDNN, random weights) dominate - the memory read is a small share here."""
    import evavos_b200 as ev
    from evavos_b200.networks import seeded_init
    _, _, t, _, _, k, seed, desc = cfg
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.set_grad_enabled(False)
    prop, fuse = ev.PropagationNetwork().eval().to(dev), ev.FusionNet().eval().to(dev)
    seeded_init(prop, 1001)
    seeded_init(fuse, 1002)
    h, w = 480, 854
    g = torch.Generator().manual_seed(seed + rank)
    videos = [torch.rand(1, t, 3, h, w, generator=g) for _ in range(2)]
    mask = (torch.rand(1, 1, h // 8, (w + 7) // 8, generator=g) > 0.6).float()
    mask = mask.repeat_interleave(8, 2).repeat_interleave(8, 3)[:, :, :h, :w]

    spread = []     # per-video wall-clock spread of every measured engine (the headline `value` is the mean)

    def measure(**opts):
        def one_video(i):
            proc = ev.InferenceCore(prop, fuse, videos[i % 2], k, device=dev, **opts)
            return proc.interact(mask, 0)
        for i in range(max(2, min(args.warmup, 3))):     # (graph capture, cuDNN autotuning and pinned staging warm up here)
            one_video(i)
        torch.cuda.synchronize(dev)
        per_video = []
        c0 = time.perf_counter()
        for i in range(args.steps):
            v0 = time.perf_counter()
            out = one_video(i)                 # (returns host masks: every video ends with a device -> host read)
            per_video.append(time.perf_counter() - v0)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - c0
        spread.append({"opts": opts, "median_ms": 1e3 * sorted(per_video)[len(per_video) // 2], "min_ms": 1e3 * min(per_video),
                       "max_ms": 1e3 * max(per_video)})
        return world * args.steps * (t - 1) / dt, 1e3 * dt / args.steps, out

    # the engine as shipped: fp32 storage, TF32 convolutions (PyTorch's default, i.e. the reference's arithmetic on this
    # GPU), BatchNorm folded into fused cuDNN conv-bias-ReLU calls, NHWC, conv passes replayed from CUDA graphs
    fps, ms, out = measure(channels_last=True, cuda_graphs=True)
    plain_fps, plain_ms, _ = measure(fold_bn=False)                      # the same engine without the conv rewrites
    amp_fps, amp_ms, _ = measure(amp=True, cuda_graphs=True)             # bf16 autocast variant (reduced precision)
    untail_fps, untail_ms, _ = measure(channels_last=True, cuda_graphs=True, fused_tails=False)   # decoder tails in PyTorch
    if rank != 0:
        return
    # the reference's stock loop on this GPU (oracle/stock_engine.py: frame by frame, dense affinity + topk + scatter,
    # bmm per object, modules as they are) - what `InferenceCore.interact` of the reference does here before switching
    try:
        from oracle.stock_engine import StockEngine
        n_ref = max(1, min(args.steps, 3))
        StockEngine(prop, fuse, videos[0], k, device=dev).interact(mask, 0)
        torch.cuda.synchronize(dev)
        c0 = time.perf_counter()
        for i in range(n_ref):
            ref_masks = StockEngine(prop, fuse, videos[i % 2], k, device=dev).interact(mask, 0)
        torch.cuda.synchronize(dev)
        ref_dt = (time.perf_counter() - c0) / n_ref
        last = (args.steps - 1) % 2
        if (n_ref - 1) % 2 != last:
            ref_masks = StockEngine(prop, fuse, videos[last], k, device=dev).interact(mask, 0)
        gpu_base = {"value": (t - 1) / ref_dt, "unit": "frames/s", "ms_per_step": 1e3 * ref_dt, "videos": n_ref,
                    "kind": "port of the reference's InferenceCore loop on CUDA tensors (oracle/stock_engine.py), 1 GPU",
                    "speedup": fps / world / ((t - 1) / ref_dt),
                    "mask_agreement": float((torch.as_tensor(ref_masks) == torch.as_tensor(out).cpu()).float().mean())}
    except Exception as e:
        gpu_base = {"unavailable": repr(e)[:200]}
    from evavos_b200.conv_opt import conv_passes
    line = {
        "metric": "propagated frames/sec (end-to-end interact)", "value": fps, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (cuDNN TF32 convolutions, fp32 keys / values / memory read)",
        "data": "synthetic", "config": {"workload": args.workload + ": " + desc, "frames": t, "image": [h, w], "mem_freq": 5,
                                         "engine": "fold_bn + fused conv-bias-ReLU, channels_last, cuda_graphs, fused decoder tails",
                                         "fused_conv_ops": bool(conv_passes(prop, False, True, True, True).fused),
                                         "note": "wall clock incl. H2D of the video and D2H of the masks; random weights "
                                                 "(every candidate list overflows: the exact tiled pass runs on every read)"},
        "mask_shape": list(out.shape),
        "plain_engine": {"value": plain_fps, "ms_per_step": plain_ms,
                         "note": "fold_bn=False, NCHW, no graphs: PyTorch modules as they are around the same memory read"},
        "gpu_baseline": gpu_base,
        "per_video_ms": spread,
        "without_fused_decoder_tails": {"value": untail_fps, "ms_per_step": untail_ms},
        "amp_variant": {"value": amp_fps, "ms_per_step": amp_ms, "dtype": "bf16 autocast convolutions (reduced precision: "
                        "not the headline), fp32 keys / values / memory read"},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- the headline arm
def bind_host_to_gpu(index):
    """Pin this process (and so its pinned staging buffers: first touch) to the CPUs NVML lists as local to GPU `index`.
    With one process per GPU streaming ~70 GB/s of PCIe traffic, buffers on the other socket cost the end-to-end leg
    its scaling.  Returns a short description for the JSON line; never fails the run."""
    if os.environ.get("EVAVOS_BIND", "1") == "0":
        return "off"
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = {64 * wi + b for wi, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return f"no-op ({len(allowed)} cpus allowed, all local)"
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} of {len(allowed)} cpus"
    except Exception as e:      # no NVML, containers without topology, ...
        return f"unavailable ({type(e).__name__})"


def run_ours(args, cfg, rank, world, local_rank):
    import evavos_b200 as ev
    from evavos_b200.host_api import memory_read_host

    ck, cv, t, h, w, k, seed, desc = cfg
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    host_binding = bind_host_to_gpu(local_rank) if world > 1 else "single process"
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    hw, n_pos = h * w, t * h * w
    pk = peaks()
    # --- resident inputs: N_BANKS distinct banks (rank-specific seeds: independent videos per GPU) ---
    banks, queries = build_banks(cfg, dev, N_BANKS, seed_shift=10007 * rank)
    q1 = [q[:, :, 0].contiguous() for q in queries]
    prob = torch.rand(k, 1, h * 16, w * 16, device=dev)
    stream = torch.cuda.current_stream(dev)
    outs = [torch.empty((k, cv, h, w), dtype=torch.float32, device=dev) for _ in range(2)]
    outs5 = [torch.empty((k, cv, MEM_FREQ, h, w), dtype=torch.float32, device=dev) for _ in range(2)]

    def read(i):
        return ev.memory_read(banks[i % N_BANKS], q1[i % N_BANKS], TOP_K, out=outs[i % 2])

    def agg(i):
        return ev.aggregate_wbg(prob, keep_bg=True)

    def step(i):
        read(i)
        return agg(i)

    def read5(i):
        return ev.memory_read(banks[i % N_BANKS], queries[i % N_BANKS], TOP_K, out=outs5[i % 2])

    def step5(i):
        read5(i)
        for _ in range(MEM_FREQ):
            agg(i)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize(dev)
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize(dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for i in range(args.steps):
        step(i)
    t1.record(stream)
    torch.cuda.synchronize(dev)
    if dist:
        dist.barrier()
    elapsed_ms = t0.elapsed_time(t1)
    clocks = sampler.stop() if rank == 0 else None
    ms5 = time_loop(step5, max(3, args.steps // MEM_FREQ), 2, stream, dev)
    if dist:
        tm = torch.tensor([elapsed_ms, ms5], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        elapsed_ms, ms5 = float(tm[0].item()), float(tm[1].item())
    steps5 = max(3, args.steps // MEM_FREQ)

    # --- instrumented pass over the same steps: per-kernel times; distinct value rows of a step ---
    stages = stage_pass(read, agg, args.steps, stream, dev)
    stages5 = stage_pass(read5, None, steps5, stream, dev)
    _, aff = ev.memory_read(banks[0], q1[0], TOP_K, want_readout=False, want_topk=True)
    distinct = int(torch.unique(aff.idx).numel())
    _, aff5 = ev.memory_read(banks[0], queries[0], TOP_K, want_readout=False, want_topk=True)
    distinct5 = int(torch.unique(aff5.idx).numel())

    # --- e2e: host buffers; streaming (bank persists) and stateless (evavos_memread_host) ---
    mk, qk, mv = synth(seed + 10007 * rank, ck, cv, t, h, w, k)
    h_mk = mk.reshape(ck, n_pos).contiguous().pin_memory()
    h_qk = qk.reshape(ck, hw).contiguous().pin_memory()
    h_mv = mv.reshape(k, cv, n_pos).contiguous().pin_memory()
    h_out = torch.empty((k, cv, hw), dtype=torch.float32).pin_memory()
    h_prob = prob.cpu().pin_memory()
    h_agg = torch.empty((k + 1, 1, h * 16, w * 16), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    s_bank = ev.MemoryBank(k, ck, cv, h, w, t, dev)              # reference-layout tensors + shadow, like do_pass
    s_bank.write_frames(0, mk.to(dev), mv.to(dev))
    h_q = [torch.randn(1, ck, h, w, generator=torch.Generator().manual_seed(seed + 7 * j)).pin_memory() for j in range(4)]
    h_newk = mk[:, :, 0].contiguous().pin_memory()
    h_newv = mv[:, :, 0:1].contiguous().pin_memory()
    s_steps = max(10, min(args.steps, 100))

    def stream_step(i):
        q = h_q[i % 4].to(dev, non_blocking=True)
        if i % MEM_FREQ == 0:    # a new memory frame every mem_freq-th step, rewritten in place in a rotating slot
            s_bank.write_frames((i // MEM_FREQ) % t, h_newk.to(dev, non_blocking=True), h_newv.to(dev, non_blocking=True))
        out, _ = ev.memory_read(s_bank, q, TOP_K)
        a = ev.aggregate_wbg(h_prob.to(dev, non_blocking=True), keep_bg=True)
        h_out.copy_(out.view(k, cv, hw), non_blocking=True)
        h_agg.copy_(a, non_blocking=True)
        torch.cuda.synchronize(dev)

    s_h2d = h_q[0].numel() * 4 + (h_newk.numel() + h_newv.numel()) * 4 // MEM_FREQ + h_prob.numel() * 4
    s_d2h = h_out.numel() * 4 + h_agg.numel() * 4
    for i in range(5):
        stream_step(i)
    if dist:
        dist.barrier()
    torch.cuda.synchronize(dev)
    c0 = time.perf_counter()
    for i in range(s_steps):
        stream_step(i)
    torch.cuda.synchronize(dev)
    stream_s = time.perf_counter() - c0

    # The same stream of steps, pipelined the way a streaming caller would run it: the uploads of step i + 1, the
    # kernels of step i and the downloads of step i - 1 overlap on three streams (PCIe is full duplex), two sets of
    # device inputs and pinned host outputs; the host consumes a step's result one step later.  Every byte still
    # crosses PCIe inside the timed region, every step.
    s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    d_q = [torch.empty((1, ck, h, w), device=dev) for _ in range(2)]
    d_p = [torch.empty_like(prob) for _ in range(2)]
    d_nk, d_nv = torch.empty((1, ck, h, w), device=dev), torch.empty((k, cv, 1, h, w), device=dev)
    d_out = [torch.empty((k, cv, h, w), device=dev) for _ in range(2)]
    h_outs = [torch.empty((k, cv, hw), dtype=torch.float32).pin_memory() for _ in range(2)]
    h_aggs = [torch.empty((k + 1, 1, h * 16, w * 16), dtype=torch.float32).pin_memory() for _ in range(2)]
    up_done = [torch.cuda.Event() for _ in range(2)]
    comp_done = [torch.cuda.Event() for _ in range(2)]
    down_done = [torch.cuda.Event() for _ in range(2)]
    aggs = [None, None]

    def upload(i):
        b = i % 2
        with torch.cuda.stream(s_up):
            s_up.wait_event(comp_done[b])            # the kernels that read this input set two steps ago are done
            d_q[b].copy_(h_q[i % 4], non_blocking=True)
            d_p[b].copy_(h_prob, non_blocking=True)
            if i % MEM_FREQ == 0:
                d_nk.copy_(h_newk, non_blocking=True)
                d_nv.copy_(h_newv, non_blocking=True)
            up_done[b].record(s_up)

    def pipelined(n):
        upload(0)
        for i in range(n):
            b = i % 2
            if i + 1 < n:
                upload(i + 1)
            stream.wait_event(up_done[b])
            stream.wait_event(down_done[b])          # the downloads that read this output set two steps ago are done
            if i % MEM_FREQ == 0:
                s_bank.write_frames((i // MEM_FREQ) % t, d_nk, d_nv)
            ev.memory_read(s_bank, d_q[b], TOP_K, out=d_out[b])
            aggs[b] = ev.aggregate_wbg(d_p[b], keep_bg=True)
            comp_done[b].record(stream)
            with torch.cuda.stream(s_down):
                s_down.wait_event(comp_done[b])
                h_outs[b].copy_(d_out[b].view(k, cv, hw), non_blocking=True)
                h_aggs[b].copy_(aggs[b], non_blocking=True)
                aggs[b].record_stream(s_down)
                down_done[b].record(s_down)
            if i >= 1:
                down_done[1 - b].synchronize()       # the host holds the result of step i - 1 now
        down_done[(n - 1) % 2].synchronize()

    pipelined(6)
    if dist:
        dist.barrier()
    torch.cuda.synchronize(dev)
    c0 = time.perf_counter()
    pipelined(s_steps)
    torch.cuda.synchronize(dev)
    pipe_s = time.perf_counter() - c0

    def e2e_step():
        h2d, d2h = memory_read_host(h_mk, h_qk, h_mv, TOP_K, out=h_out)
        p = h_prob.to(dev, non_blocking=True)
        h_agg.copy_(ev.aggregate_wbg(p, keep_bg=True), non_blocking=True)
        torch.cuda.synchronize(dev)
        return h2d + h_prob.numel() * 4, d2h + h_agg.numel() * 4

    for _ in range(3):
        h2d_b, d2h_b = e2e_step()
    if dist:
        dist.barrier()
    torch.cuda.synchronize(dev)
    c0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - c0
    if dist:
        tm = torch.tensor([stream_s, e2e_s, pipe_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        stream_s, e2e_s, pipe_s = float(tm[0].item()), float(tm[1].item()), float(tm[2].item())
        dist.barrier()

    # --- the reference's own torch sequence on this GPU (rank 0's number is reported) ---
    gb = gpu_baseline(cfg, (mk.to(dev), qk.to(dev), mv.to(dev)), max(3, min(args.steps, 10)), dev, stream)
    del s_bank
    torch.cuda.empty_cache()

    # --- N > 1: the memory-axis sharded long-video read (BASELINE.json configs[3]) over the same process group ---
    sharded = None
    if world > 1:
        sharded = run_sharded(args, WORKLOADS["cfg4"], rank, world, local_rank, dist=dist, emit=False)
        # the same read with the queries split as well: world / M query groups x M memory shards (1 / M of the bank per GPU)
        hybrid = {}
        for m in (world // 2, world // 4, 1):
            if m >= 1 and world % m == 0 and f"x{m}" not in hybrid:
                r = run_sharded(args, WORKLOADS["cfg4"], rank, world, local_rank, dist=dist, emit=False, memory_shards=m)
                if r is not None:
                    hybrid[f"x{m}"] = {key: r[key] for key in ("value", "ms_per_step", "efficiency_vs_same_run_single_gpu",
                                                              "speedup_vs_same_run_single_gpu", "per_rank_stage_us_max",
                                                              "parity_ok", "parity_rel_l2_max")}
                    hybrid[f"x{m}"].update(parallelism=r["config"]["parallelism"], bank_fraction_per_gpu=1.0 / m)
                else:
                    hybrid[f"x{m}"] = None
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    step_us = 1e3 * elapsed_ms / args.steps
    shape = (ck, cv, n_pos, hw, k, h * 16 * w * 16)
    dom, kernels, step_roof = rooflines(shape, stages, step_us, distinct, 4, pk)
    dom5, kernels5, step_roof5 = rooflines(shape, dict(stages5), 1e3 * ms5 / steps5, distinct5, 4, pk, frames=MEM_FREQ)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        traffic = tr.get(args.workload, {}).get(dom["kernel"])
        dom["traffic_source"] = tr.get("source")
    except Exception:
        pass
    value = world * args.steps / (elapsed_ms * 1e-3)
    line = {
        "metric": "memory-read query-frames/sec", "value": value, "unit": "query-frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + desc, "top_k": TOP_K, "CK": ck, "CV": cv, "memory_positions": n_pos,
                   "queries_per_frame": hw, "objects": k,
                   "l2": f"{N_BANKS} rotating banks, {N_BANKS * (4 * k * cv * n_pos + 4 * ck * n_pos) / 1e6:.0f} MB of inputs > 126 MB L2",
                   "filter": "tcgen05 bf16 candidate filter (sampled threshold pass + one candidate sweep) + exact fp32 rescoring",
                   "parallelism": f"independent videos x{world}"},
        "clocks": clocks,
        "e2e": {"value": world * s_steps / min(pipe_s, stream_s), "unit": "query-frames/s", "h2d_bytes_per_step": int(s_h2d),
                "d2h_bytes_per_step": int(s_d2h), "steps": s_steps,
                "mode": "pipelined" if pipe_s < stream_s else "synchronised every step",
                "note": "public API, pinned host buffers, every step: H2D query key + decoder probabilities + (every 5th "
                        "step) one new memory frame rewritten in place; D2H readout + aggregated probabilities; the bank "
                        "itself is engine state, as in the reference.  Pipelined = uploads of step i+1, kernels of step i "
                        "and downloads of step i-1 overlap on three streams, the host reads a result one step later"},
        "host_binding": host_binding,
        "e2e_pipelined": {"value": world * s_steps / pipe_s, "unit": "query-frames/s"},
        "e2e_sync_every_step": {"value": world * s_steps / stream_s, "unit": "query-frames/s"},
        "e2e_full_upload": {"value": world * e2e_steps / e2e_s, "unit": "query-frames/s", "h2d_bytes_per_step": int(h2d_b),
                            "d2h_bytes_per_step": int(d2h_b), "steps": e2e_steps,
                            "note": "evavos_memread_host: stateless, the whole bank + query H2D, shadow build, read, D2H, every step"},
        "gpu_launches": KERNELS_PER_STEP * args.steps,
        "stages_us": stages,
        "roofline": dict(dom, traffic=traffic, peak_source=pk["source"], kernels=kernels, step=step_roof,
                         share_of_step=dom["us_per_launch"] / sum(v["us_per_launch"] for v in kernels.values()),
                         note="dominant = longest kernel of the step; kernel times from the instrumented pass (events "
                              "between kernels defeat PDL overlap: they add up to more than ms_per_step)"),
        "batched_read": {"value": world * MEM_FREQ * steps5 / (ms5 * 1e-3), "unit": "query-frames/s",
                         "query_frames_per_launch": MEM_FREQ, "ms_per_launch": ms5 / steps5, "steps": steps5,
                         "stages_us": stages5,
                         "roofline": {"dominant": dom5["kernel"], "kernels": kernels5, "step": step_roof5},
                         "note": "the read the product issues (inference_core.py): the mem_freq query frames between two "
                                 "memory appends in ONE launch + their 5 aggregations"},
        "gpu_baseline": dict(gb, speedup=value / world / gb["value"]),
        "parity_note": "GPU tests (-m gpu) hold the gates: top-k sets vs the fp64 oracle, readout rel-L2, goldens of the live "
                       "reference; e2e goldens are 6-7 frames at <=120x150 with random weights and cuDNN TF32 off",
    }
    if world == 1:
        rate, info = cpu_reference_rate(cfg, 3, 1, budget_s=30.0)
        line["cpu_baseline"] = {"value": rate, "unit": "query-frames/s", "cores": info["cores"], "kind": "port",
                                "sample": info["sample"]}
        for name in ("cfg4", "cfg5"):
            try:
                line[name] = side_workload(name, args, dev, stream, pk)
            except Exception as e:
                line[name] = {"error": repr(e)[:300]}
        try:
            line["attention_read"] = attention_workload(dev, stream, pk)
        except Exception as e:
            line["attention_read"] = {"error": repr(e)[:300]}
    if sharded is not None:
        line["sharded_cfg4"] = sharded
        line["hybrid_cfg4"] = {k_: v_ for k_, v_ in hybrid.items() if v_ is not None}
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else max(1, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
    elif args.workload == "cfg4":
        run_sharded(args, cfg, rank, world, local_rank)
    elif args.workload == "cfg3":
        run_cfg3(args, cfg, rank, world, local_rank)
    elif args.workload == "cfg5":
        dev = torch.device("cuda", local_rank)
        torch.cuda.set_device(dev)
        if rank == 0:
            res = side_workload("cfg5", args, dev, torch.cuda.current_stream(dev), peaks())
            print(json.dumps(res), flush=True)
    else:
        run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
