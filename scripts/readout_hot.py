"""Readout time when every query selects the same rows (what near-constant keys produce) against random rows.

    python scripts/readout_hot.py
480p map, 8 100 queries (5 frames), K = 1, CV = 512, 8 100 / 32 400 positions.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from evavos_b200.memory_reader import TopKAffinity  # noqa: E402

dev = torch.device("cuda:0")
h, w, cv, k = 30, 54, 512, 50
nq = 5 * h * w
for t in (5, 20):
    n = t * h * w
    g = torch.Generator().manual_seed(t)
    bank = ev.MemoryBank(1, 64, cv, h, w, t, dev, keep_reference_layout=False)
    bank.write_frames(0, torch.randn(1, 64, t, h, w, generator=g).to(dev), torch.randn(1, cv, t, h, w, generator=g).to(dev))
    reader = ev.EvalMemoryReader(k, None)
    wgt = torch.softmax(torch.randn(nq, k, generator=g), 1).to(dev)
    cases = {
        "random rows": torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(nq)]).to(torch.int32),
        "same 50 rows for all": torch.randperm(n, generator=g)[:k].to(torch.int32).expand(nq, k).contiguous(),
        "same rows within 8 queries": torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(nq // 8 + 1)])
        .to(torch.int32).repeat_interleave(8, 0)[:nq].contiguous(),
        "neighbours share 80%": None,
    }
    base = cases["random rows"].clone()
    for q in range(1, nq):
        if q % 8:
            base[q, :40] = base[q - 1, :40]
    cases["neighbours share 80%"] = base
    for name, idx in cases.items():
        aff = TopKAffinity(idx.to(dev), wgt, None, n, h, w)
        for _ in range(3):
            reader.readout(aff, bank)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            reader.readout(aff, bank)
        e1.record()
        torch.cuda.synchronize()
        print(f"[{n} positions] {name:28s} {1e3 * e0.elapsed_time(e1) / 20:7.1f} us", flush=True)
