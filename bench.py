"""bench.py - memory-read query-frames/s of the B200 space-time memory read (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg4]

A step is one query frame of the hot path on synthetic inputs: key affinity + top-k softmax +
value readout for all objects (the fused read) followed by the soft aggregation across objects.
Workload (BASELINE.json configs[1]): DAVIS-17 480p, 30x54 feature map, 3 objects, 20-frame bank.

 * value  - whole-job query-frames/s with the bank resident in HBM (CUDA events, max over ranks).
            Four banks (each > L2 together) are rotated so no step finds its inputs in L2.
 * e2e    - the same step through the public API with pinned HOST buffers, synchronised every step.  The bank
            is engine state (exactly as the reference keeps its bank resident in host memory between frames); a
            step's inputs are the frame's query key, the decoder probabilities and - every mem_freq-th frame -
            one new memory frame (H2D + in-place append, inference_core.py:174-177); its result is the readout
            and the aggregated probabilities (D2H).  `e2e_full_upload` is the stateless variant
            (evavos_memread_host: the whole bank crosses PCIe every step).
 * roofline - the dominant kernel (sparse readout, HBM-bound), timed with CUDA events inside the
            timed region; algorithmic bytes = s*K*CV*min(N, k*HW) + 4*K*CV*HW + 8*k*HW.
 * cpu_baseline / --impl reference - oracle/torch_port.py (op-for-op port of the reference's dense
            torch path; /root/reference is not on the GPU box) on all host cores.
Multi-GPU (torchrun, one rank per GPU): independent videos are partitioned across ranks with no
collective -> weak scaling; value = N * K / max-over-ranks time.
`--workload cfg4` is the long-video case instead: ONE 200-frame bank sharded along the memory axis over
the ranks (NCCL all-gather of top-k candidates + all-reduce of partial readouts) -> strong scaling.
"""
from __future__ import annotations

import argparse
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
    "cfg4": (64, 512, 200, 30, 54, 1, 1238, "MOSE-style long video 480p, 1 object, 200-frame bank (unsharded)"),
    "cfg3": (64, 512, 32, 30, 54, 1, 1237, "independent synthetic 480p videos (32 frames, 1 object), full key/value encode + memory read + decode"),
    "cfg5": (64, 512, 50, 68, 120, 5, 1239, "1080p-equivalent feature map (68x120), 5 objects, 50-frame bank, bf16 value bank"),
}
TOP_K = 50
N_BANKS = 4
FRAMES_PER_EXCHANGE = 5   # query frames between two memory appends (mem_freq) read in one sharded exchange
KERNELS_PER_STEP = 5  # fused score filter, exact fallback (overflow), finalize, readout, aggregate (+ one memset node)


def synth(seed, ck, cv, t, h, w, k):
    g = torch.Generator().manual_seed(seed)
    mk = torch.randn(1, ck, t, h, w, generator=g)
    qk = torch.randn(1, ck, h, w, generator=g)
    mv = torch.randn(k, cv, t, h, w, generator=g)
    return mk, qk, mv


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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


def run_sharded(args, cfg, rank, world, local_rank):
    """cfg4: one long bank sharded by frame over the ranks; strong scaling (total work fixed)."""
    import evavos_b200 as ev
    from evavos_b200.sharded import ShardedMemoryBank
    ck, cv, t, h, w, k, seed, desc = cfg
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    hw, n_pos = h * w, t * h * w
    n_banks = 2
    banks, queries = [], []
    for b in range(n_banks):
        g = torch.Generator().manual_seed(seed + 100 * b)
        bank = ShardedMemoryBank(k, ck, cv, h, w, t, dev)
        for f in range(t):                      # same frames on every rank; only the owner keeps one
            kf = torch.randn(1, ck, h, w, generator=g)
            vf = torch.randn(k, cv, 1, h, w, generator=g)
            if bank.owner_of(f) == rank:
                bank.append(kf.to(dev), vf.to(dev))
            else:
                bank.n_frames += 1
        banks.append(bank)
        queries.append(torch.randn(1, ck, FRAMES_PER_EXCHANGE, h, w, generator=g).to(dev))
    prob = torch.rand(k, 1, h * 16, w * 16, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step(i):
        # mem_freq = 5 query frames share one bank state (inference_core.py:174): one exchange for all of them
        out = banks[i % n_banks].read(queries[i % n_banks], TOP_K)
        aggs = [ev.aggregate_wbg(prob, keep_bg=True) for _ in range(FRAMES_PER_EXCHANGE)]
        return out, aggs

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
    if dist:
        tm = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tm.item())
    if rank == 0:
        value = FRAMES_PER_EXCHANGE * args.steps / (elapsed_ms * 1e-3)
        flops = 2.0 * n_pos * hw * ck * FRAMES_PER_EXCHANGE                       # the one unavoidable dense contraction
        t_tc = flops / 1390.2e12
        line = {
            "metric": "memory-read query-frames/sec", "value": value, "unit": "query-frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + desc + f", memory axis sharded by frame over {world} GPU(s)",
                       "top_k": TOP_K, "memory_positions": n_pos, "queries_per_frame": hw, "objects": k,
                       "query_frames_per_step": FRAMES_PER_EXCHANGE,
                       "l2": f"{n_banks} rotating banks, {n_banks * 4 * (k * cv + ck) * n_pos / 1e6:.0f} MB of inputs > 126 MB L2",
                       "parallelism": f"memory-axis shards x{world}, all-gather(top-k) + all-reduce(readout)"},
            "clocks": clocks, "gpu_launches": (KERNELS_PER_STEP + 1 + FRAMES_PER_EXCHANGE - 1) * args.steps,
            "roofline": {"bound": "tensor", "kernel": "whole sharded read (score filter dominates)", "achieved": flops / (elapsed_ms / args.steps * 1e-3) / 1e12 / world,
                         "peak": 1390.2, "unit": "TFLOP/s", "frac": t_tc / world / (elapsed_ms / args.steps * 1e-3), "traffic": None,
                         "note": "algorithmic flops 2*N*HW*CK per query frame over the measured step time, per GPU"},
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def run_cfg5(args, cfg, rank, world, local_rank):
    """BASELINE.json configs[4] on one GPU per rank (independent replicas): bf16-representable inputs, bf16 value
    shadow, fp32 accumulation.  Banks are synthesised frame by frame on the device (10 GB of fp32 values per
    bank would not fit a sensible host buffer); stage times come from the library's event diagnostics."""
    import ctypes
    import numpy as np
    import evavos_b200 as ev
    from evavos_b200 import _lib
    ck, cv, t, h, w, k, seed, desc = cfg
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    hw, n_pos, n_banks = h * w, t * h * w, 2
    banks, queries = [], []
    g = torch.Generator(device=dev).manual_seed(seed + rank)
    for b in range(n_banks):
        bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, value_dtype=torch.bfloat16, keep_reference_layout=False)
        for f in range(t):
            kf = torch.randn(1, ck, h, w, generator=g, device=dev).to(torch.bfloat16).float()
            vf = torch.randn(k, cv, 1, h, w, generator=g, device=dev).to(torch.bfloat16).float()
            bank.append(kf, vf)
        banks.append(bank)
        queries.append(torch.randn(1, ck, h, w, generator=g, device=dev).to(torch.bfloat16).float())
    prob = torch.rand(k, 1, h * 16, w * 16, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step(i):
        out, _ = ev.memory_read(banks[i % n_banks], queries[i % n_banks], TOP_K)
        return out, ev.aggregate_wbg(prob, keep_bg=True)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize(dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for i in range(args.steps):
        step(i)
    t1.record(stream)
    torch.cuda.synchronize(dev)
    elapsed_ms = t0.elapsed_time(t1)
    lib = _lib.load()
    lib.evavos_stage_timing(1)
    acc = np.zeros(4)
    for i in range(5):
        step(i)
        ms = (ctypes.c_float * 4)()
        lib.evavos_stage_timing_read(ms)
        acc += np.array(list(ms)) / 5
    lib.evavos_stage_timing(0)
    if rank != 0:
        return
    flops = 2.0 * n_pos * hw * ck + 2.0 * k * cv * TOP_K * hw
    bytes_alg = 2.0 * (n_pos * ck + hw * ck + k * cv * min(n_pos, TOP_K * hw)) + 4.0 * k * cv * hw
    t_roof = max(flops / 1390.2e12, bytes_alg / (peaks()[0] * 1e9))
    line = {
        "metric": "memory-read query-frames/sec", "value": world * args.steps / (elapsed_ms * 1e-3), "unit": "query-frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": args.workload + ": " + desc, "top_k": TOP_K, "memory_positions": n_pos, "queries_per_frame": hw,
                   "objects": k, "l2": f"{n_banks} rotating banks of {2 * k * cv * n_pos / 1e9:.1f} GB"},
        "gpu_launches": KERNELS_PER_STEP * args.steps,
        "stages_us": {"filter": acc[0] * 1e3, "exact_fallback": acc[1] * 1e3, "finalize": acc[2] * 1e3, "readout": acc[3] * 1e3},
        "roofline": {"bound": "balanced (SURVEY 8d: 308 us tensor vs 333 us HBM)", "achieved": None, "peak": None,
                     "unit": "fraction of max(F_alg / P_tc, B_alg / BW_hbm)", "frac": t_roof / (elapsed_ms / args.steps * 1e-3),
                     "traffic": None, "t_roof_us": t_roof * 1e6},
        "cpu_baseline": {"value": None, "unit": "query-frames/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "skipped: the dense fp32 affinity of this config is 13.3 GB (x3 temporaries)"},
    }
    print(json.dumps(line), flush=True)


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

    def one_video(i):
        proc = ev.InferenceCore(prop, fuse, videos[i % 2], k, device=dev)
        return proc.interact(mask, 0)

    for i in range(max(1, min(args.warmup, 2))):
        one_video(i)
    torch.cuda.synchronize(dev)
    c0 = time.perf_counter()
    for i in range(args.steps):
        out = one_video(i)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - c0
    if rank != 0:
        return
    line = {
        "metric": "propagated frames/sec (end-to-end interact)", "value": world * args.steps * (t - 1) / dt, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (cuDNN TF32 convolutions)",
        "data": "synthetic", "config": {"workload": args.workload + ": " + desc, "frames": t, "image": [h, w], "mem_freq": 5,
                                         "note": "wall clock incl. H2D of the video and D2H of the masks; random weights"},
        "mask_shape": list(out.shape),
    }
    print(json.dumps(line), flush=True)


def run_ours(args, cfg, rank, world, local_rank):
    import evavos_b200 as ev
    from evavos_b200 import _lib
    from evavos_b200.host_api import memory_read_host
    import ctypes

    ck, cv, t, h, w, k, seed, desc = cfg
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    hw, n_pos = h * w, t * h * w
    # --- resident inputs: N_BANKS distinct banks (rank-specific seeds: independent videos per GPU) ---
    banks, queries = [], []
    for b in range(N_BANKS):
        mk, qk, mv = synth(seed + 100 * b + 10007 * rank, ck, cv, t, h, w, k)
        bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, keep_reference_layout=False)
        for f in range(t):  # built the way do_pass builds it: one append per memory frame
            bank.append(mk[:, :, f].to(dev), mv[:, :, f:f + 1].to(dev))
        banks.append(bank)
        queries.append(qk.to(dev))
    prob = torch.rand(k, 1, h * 16, w * 16, device=dev)
    idx = torch.empty((hw, TOP_K), dtype=torch.int32, device=dev)
    wgt = torch.empty((hw, TOP_K), dtype=torch.float32, device=dev)
    lib = _lib.load()
    stream = torch.cuda.current_stream(dev)

    ro_events = []

    def step(i, timed):
        bank, qk = banks[i % N_BANKS], queries[i % N_BANKS]
        # fused read, split at the readout only to bracket the dominant kernel with events
        _, aff = ev.memory_read(bank, qk, TOP_K, want_readout=False, want_topk=True)
        out = torch.empty((k, cv, hw), dtype=torch.float32, device=dev)
        sh = bank.shadow()
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        _lib.check(lib.evavos_readout(ctypes.byref(sh), aff.idx.data_ptr(), aff.weight.data_ptr(), hw, TOP_K,
                                      out.data_ptr(), 0, 0, stream.cuda_stream))
        if timed:
            e1.record(stream)
            ro_events.append((e0, e1))
        agg = ev.aggregate_wbg(prob, keep_bg=True)
        return out, agg

    for i in range(args.warmup):
        step(i, False)
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
        step(i, True)
    t1.record(stream)
    torch.cuda.synchronize(dev)
    if dist:
        dist.barrier()
    elapsed_ms = t0.elapsed_time(t1)
    clocks = sampler.stop() if rank == 0 else None
    ro_ms = sum(a.elapsed_time(b) for a, b in ro_events) / len(ro_events)
    if dist:
        tm = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tm.item())

    # --- e2e: host buffers through the host-buffer C-ABI entry point, fewer steps (PCIe-bound) ---
    mk, qk, mv = synth(seed + 10007 * rank, ck, cv, t, h, w, k)
    h_mk = mk.reshape(ck, n_pos).contiguous().pin_memory()
    h_qk = qk.reshape(ck, hw).contiguous().pin_memory()
    h_mv = mv.reshape(k, cv, n_pos).contiguous().pin_memory()
    h_out = torch.empty((k, cv, hw), dtype=torch.float32).pin_memory()
    h_prob = prob.cpu().pin_memory()
    h_agg = torch.empty((k + 1, 1, h * 16, w * 16), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))

    # streaming e2e: the bank persists; per step H2D = query (+ 1/mem_freq of a new memory frame) + probabilities
    mem_freq = 5
    s_bank = ev.MemoryBank(k, ck, cv, h, w, t, dev)              # reference-layout tensors + shadow, like do_pass
    s_bank.write_frames(0, mk.to(dev), mv.to(dev))
    h_q = [torch.randn(1, ck, h, w, generator=torch.Generator().manual_seed(seed + 7 * j)).pin_memory() for j in range(4)]
    h_newk = mk[:, :, 0].contiguous().pin_memory()
    h_newv = mv[:, :, 0:1].contiguous().pin_memory()
    s_steps = max(10, min(args.steps, 100))

    def stream_step(i):
        q = h_q[i % 4].to(dev, non_blocking=True)
        if i % mem_freq == 0:
            s_bank.write_frames((i // mem_freq) % t, h_newk.to(dev, non_blocking=True), h_newv.to(dev, non_blocking=True))
        out, _ = ev.memory_read(s_bank, q, TOP_K)
        agg = ev.aggregate_wbg(h_prob.to(dev, non_blocking=True), keep_bg=True)
        h_out.copy_(out.view(k, cv, hw), non_blocking=True)
        h_agg.copy_(agg, non_blocking=True)
        torch.cuda.synchronize(dev)

    s_h2d = h_q[0].numel() * 4 + (h_newk.numel() + h_newv.numel()) * 4 // mem_freq + h_prob.numel() * 4
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
    if dist:
        tm = torch.tensor([stream_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        stream_s = float(tm.item())

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
        tm = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = float(tm.item())
        dist.barrier()
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    value = world * args.steps / (elapsed_ms * 1e-3)
    peak, peak_src = peaks()
    # readout kernel, per launch: every needed value row once + output once + (idx, weight) once
    ro_bytes = 4 * k * cv * min(n_pos, TOP_K * hw) + 4 * k * cv * hw + 8 * TOP_K * hw
    achieved = ro_bytes / (ro_ms * 1e-3) / 1e9
    line = {
        "metric": "memory-read query-frames/sec", "value": value, "unit": "query-frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + desc, "top_k": TOP_K, "CK": ck, "CV": cv, "memory_positions": n_pos,
                   "queries_per_frame": hw, "objects": k,
                   "l2": f"{N_BANKS} rotating banks, {N_BANKS * (4 * k * cv * n_pos + 4 * ck * n_pos) / 1e6:.0f} MB of inputs > 126 MB L2",
                   "filter": "tcgen05 bf16 candidate filter + exact fp32 rescoring", "parallelism": f"independent videos x{world}"},
        "clocks": clocks,
        "e2e": {"value": world * s_steps / stream_s, "unit": "query-frames/s", "h2d_bytes_per_step": int(s_h2d),
                "d2h_bytes_per_step": int(s_d2h), "steps": s_steps,
                "note": "public API, pinned host buffers, synchronised every step: H2D query key + decoder probabilities + "
                        "(every 5th step) one new memory frame appended in place; D2H readout + aggregated probabilities; "
                        "the bank itself is engine state, as in the reference.  (A two-frames-in-flight variant with the "
                        "copies on their own streams was measured slower on this platform: 484 vs 1047 qf/s.)"},
        "e2e_full_upload": {"value": world * e2e_steps / e2e_s, "unit": "query-frames/s", "h2d_bytes_per_step": int(h2d_b),
                            "d2h_bytes_per_step": int(d2h_b), "steps": e2e_steps,
                            "note": "evavos_memread_host: stateless, the whole bank + query H2D, shadow build, read, D2H, every step"},
        "gpu_launches": KERNELS_PER_STEP * args.steps,
        "roofline": {"bound": "hbm", "kernel": "readout_f32_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "us_per_launch": ro_ms * 1e3, "algorithmic_bytes": ro_bytes,
                     "share_of_step": ro_ms / (elapsed_ms / args.steps)},
    }
    traffic_file = os.path.join(ROOT, "profiles", "readout_traffic.json")
    if os.path.exists(traffic_file):
        try:
            line["roofline"]["traffic"] = json.load(open(traffic_file)).get(args.workload)
        except Exception:
            pass
    if world == 1:
        cpu_steps = 3
        rate, info = cpu_reference_rate(cfg, cpu_steps, 1, budget_s=30.0)
        line["cpu_baseline"] = {"value": rate, "unit": "query-frames/s", "cores": info["cores"], "kind": "port",
                                "sample": info["sample"]}
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
    elif args.workload == "cfg5":
        run_cfg5(args, cfg, rank, world, local_rank)
    elif args.workload == "cfg3":
        run_cfg3(args, cfg, rank, world, local_rank)
    else:
        run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
