"""Stress the fused score filter: many launches over varied bank / query shapes, tensor path vs exact SIMT path.

    timeout 200 python scripts/stress_filter.py [rounds]
Guards against timing-dependent hangs of the cooperative kernel (run it under `timeout`) and against any
difference between the two selection paths.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from evavos_b200 import _lib  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(2024)
shapes = [(30, 54), (9, 13), (17, 31), (68, 120), (1, 50), (12, 11)]
launches, t_start = 0, time.time()
for r in range(rounds):
    h, w = shapes[r % len(shapes)]
    t = int(torch.randint(1, 40 if h * w < 2000 else 8, (1,), generator=g))
    if t * h * w < 50:
        t = 50 // (h * w) + 1
    k = int(torch.randint(1, 4, (1,), generator=g))
    f = int(torch.randint(1, 4, (1,), generator=g))
    mk = torch.randn(1, 64, t, h, w, generator=g).to(dev)
    mv = torch.randn(k, 512, t, h, w, generator=g).to(dev)
    qk = (torch.randn(1, 64, f, h, w, generator=g) if f > 1 else torch.randn(1, 64, h, w, generator=g)).to(dev)
    bank = ev.MemoryBank.from_tensors(mk, mv)
    ref_out, ref_aff = ev.memory_read(bank, qk, 50, want_topk=True, path=_lib.PATH_SIMT)
    for rep in range(30):
        out, aff = ev.memory_read(bank, qk, 50, want_topk=True, path=_lib.PATH_TENSOR)
        launches += 1
        if rep % 10 == 0:
            torch.cuda.synchronize()
            assert torch.equal(aff.idx, ref_aff.idx), f"round {r}: index mismatch (T={t}, {h}x{w}, F={f})"
            assert (out - ref_out).abs().max().item() < 1e-5
    torch.cuda.synchronize()
    del bank
print(f"stress ok: {launches} filter launches over {rounds} shapes in {time.time() - t_start:.1f} s", flush=True)
