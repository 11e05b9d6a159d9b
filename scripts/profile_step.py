"""A few steps of each workload for ncu captures (round 2): every kernel of the library is launched at least twice.

    python scripts/profile_step.py [cfg2 cfg4 cfg5 extras]
cfg2 / cfg4 / cfg5: bank append (write_keys / write_values), fused read (score_select, finalize, readout), aggregate.
extras: attention read at 68x120, sharded merge on gathered lists, J&F metric at 480x854, argmax/unpad.
dense: near-constant keys, every list overflows -> overflow_exact_kernel.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from bench import TOP_K, WORKLOADS  # noqa: E402
from evavos_b200.sharded import CudaShardOps  # noqa: E402

dev = torch.device("cuda:0")
names = sys.argv[1:] or ["cfg2", "cfg4", "cfg5", "extras"]
for name in names:
    if name == "extras":
        g = torch.Generator(device=dev).manual_seed(1)
        mk = torch.randn(1, 64, 68, 120, generator=g, device=dev)
        qk = torch.randn(1, 64, 68, 120, generator=g, device=dev)
        vec = torch.rand(10, 68 * 120, generator=g, device=dev)
        for _ in range(2):
            ev.attention_readout(mk, qk, vec)
        ops = CudaShardOps()
        nq, world = 8100, 8
        gathered = torch.stack([torch.randint(0, 25 * 1620, (world, nq, TOP_K), generator=g, device=dev).int(),
                                torch.randn(world, nq, TOP_K, generator=g, device=dev).sort(-1, descending=True).values.view(torch.int32)], -1).contiguous()
        for _ in range(2):
            ops.merge_gathered(gathered, TOP_K, 0, world, 1620)
        pred = torch.rand(32, 480, 854, generator=g, device=dev) > 0.5
        gt = torch.rand(32, 480, 854, generator=g, device=dev) > 0.5
        for _ in range(2):
            ev.frame_metrics(pred, gt)
        prob = torch.rand(4, 32, 1, 480, 864, generator=g, device=dev)
        for _ in range(2):
            ev.argmax_unpad(prob, (5, 5, 0, 0), 480, 854)
        torch.cuda.synchronize()
        continue
    if name == "dense":
        # near-constant keys (what random-weight networks produce): every candidate list overflows and the exact tiled
        # pass (overflow_exact_kernel) selects for all 8 100 queries of a 5-frame read against a 5-frame bank
        from evavos_b200 import _lib
        g = torch.Generator(device=dev).manual_seed(5)
        t, h, w = 5, 30, 54
        base = torch.randn(1, 64, 1, 1, 1, generator=g, device=dev)
        bank = ev.MemoryBank(1, 64, 512, h, w, t, dev, keep_reference_layout=False)
        bank.write_frames(0, base + 1e-3 * torch.randn(1, 64, t, h, w, generator=g, device=dev),
                          torch.randn(1, 512, t, h, w, generator=g, device=dev))
        qk = 0.9 * base + 1e-3 * torch.randn(1, 64, 5, h, w, generator=g, device=dev)
        for _ in range(3):
            ev.memory_read(bank, qk, TOP_K, path=_lib.PATH_TENSOR_DENSE)
        torch.cuda.synchronize()
        print(name, "done", flush=True)
        continue
    ck, cv, t, h, w, k, seed, _ = WORKLOADS[name]
    bf16 = name == "cfg5"
    g = torch.Generator(device=dev).manual_seed(seed)
    bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, value_dtype=torch.bfloat16 if bf16 else torch.float32,
                         keep_reference_layout=not bf16)
    # the whole bank in a few launches (under ncu every profiled launch is replayed ~40 times: 200 appends would
    # take longer than everything else), then two single-frame appends so that the append kernels are captured too
    step = max(1, t // 4)
    for f0 in range(0, t - 2, step):
        f1 = min(t - 2, f0 + step)
        bank.write_frames(f0, torch.randn(1, ck, f1 - f0, h, w, generator=g, device=dev),
                          torch.randn(k, cv, f1 - f0, h, w, generator=g, device=dev))
    for f in range(t - 2, t):
        bank.append(torch.randn(1, ck, h, w, generator=g, device=dev), torch.randn(k, cv, 1, h, w, generator=g, device=dev))
    qk = torch.randn(1, ck, h, w, generator=g, device=dev)
    prob = torch.rand(k, 1, h * 16, w * 16, generator=g, device=dev)
    for _ in range(4):
        ev.memory_read(bank, qk, TOP_K)
        ev.aggregate_wbg(prob, keep_bg=True)
    torch.cuda.synchronize()
    print(name, "done", flush=True)
    del bank
    torch.cuda.empty_cache()
