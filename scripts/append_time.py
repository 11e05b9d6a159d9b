"""Time of one memory-bank append (keys + values of one frame: reference layout + position-major shadow)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402

dev = torch.device("cuda:0")
for (k, h, w, bf16) in [(3, 30, 54, False), (1, 30, 54, False), (5, 68, 120, True)]:
    t = 24
    bank = ev.MemoryBank(k, 64, 512, h, w, t, dev, value_dtype=torch.bfloat16 if bf16 else torch.float32,
                         keep_reference_layout=not bf16)
    kf = torch.randn(1, 64, h, w, device=dev)
    vf = torch.randn(k, 512, 1, h, w, device=dev)
    for i in range(4):
        bank.write_frames(i, kf.unsqueeze(2), vf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        bank.write_frames(4 + i, kf.unsqueeze(2), vf)
    e1.record()
    torch.cuda.synchronize()
    mb = (64 + k * 512) * h * w * 4 * (3 if not bf16 else 1.5) / 1e6
    us = 1e3 * e0.elapsed_time(e1) / 20
    print(f"K={k} {h}x{w} {'bf16' if bf16 else 'fp32'} values: append {us:.1f} us ({mb:.1f} MB moved, {mb / us:.2f} TB/s)", flush=True)
