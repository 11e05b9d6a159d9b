"""Minimal driver for ncu: one cfg bank, a few fused reads + aggregate (used to capture kernel profiles)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from bench import WORKLOADS, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ck, cv, t, h, w, k, seed, _ = WORKLOADS[name]
dev = torch.device("cuda:0")
mk, qk, mv = synth(seed, ck, cv, t, h, w, k)
bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, keep_reference_layout=False)
bank.write_frames(0, mk.to(dev), mv.to(dev))
qk = qk.to(dev)
prob = torch.rand(k, 1, h * 16, w * 16, device=dev)
for _ in range(iters):
    out, _ = ev.memory_read(bank, qk, 50)
    agg = ev.aggregate_wbg(prob, keep_bg=True)
torch.cuda.synchronize()
print("done", float(out.abs().mean()), float(agg.sum()))
