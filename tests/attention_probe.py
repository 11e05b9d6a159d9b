"""Time the fused attention read against the torch op sequence it replaces (prop_net.py:117-138, 204-207)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from oracle import torch_port as port  # noqa: E402  (timed as "what the reference does on this GPU", not shipped)


def timed(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters * 1e3


for h, w, b in ((30, 54, 2), (30, 54, 4), (68, 120, 6)):
    g = torch.Generator().manual_seed(1)
    mk = torch.randn(1, 64, 1, h, w, generator=g).cuda()
    qk = torch.randn(1, 64, h, w, generator=g).cuda()
    vec = torch.rand(2 * b, h * w, generator=g).cuda()
    us = timed(lambda: ev.attention_readout(mk, qk, vec))
    us_ref = timed(lambda: vec.view(2 * b, 1, -1) @ port.attention_weights(mk, qk))
    print(f"[{h}x{w}, {2 * b} rows] fused attention read {us:.1f} us | torch dense W + matmul {us_ref:.1f} us "
          f"(W = {4 * (h * w) ** 2 / 1e6:.1f} MB)", flush=True)
