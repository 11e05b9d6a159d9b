"""cfg2 bank, mem_freq = 5 query frames per launch (the dense-hit regime of the filter) - for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from bench import TOP_K, WORKLOADS  # noqa: E402

dev = torch.device("cuda:0")
ck, cv, t, h, w, k, seed, _ = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
g = torch.Generator(device=dev).manual_seed(seed)
bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, keep_reference_layout=False)
bank.write_frames(0, torch.randn(1, ck, t, h, w, generator=g, device=dev), torch.randn(k, cv, t, h, w, generator=g, device=dev))
qk = torch.randn(1, ck, 5, h, w, generator=g, device=dev)
for _ in range(4):
    ev.memory_read(bank, qk, TOP_K)
torch.cuda.synchronize()
