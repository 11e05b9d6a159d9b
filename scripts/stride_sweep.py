"""Filter time against the sample stride R of the threshold pass, over bank lengths and query-frame counts.

    python scripts/stride_sweep.py
480p feature maps (30x54), K = 1; prints filter / finalize us for R = 1, 2 (the data behind score_pass_sample_stride).
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from evavos_b200 import _lib  # noqa: E402
from bench import synth  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
h, w = 30, 54
for t, f in [(20, 1), (20, 5), (25, 5), (50, 1), (50, 5), (100, 1), (100, 5), (200, 1), (200, 5)]:
    mk, _, mv = synth(11 + t, 64, 512, t, h, w, 1)
    qk = torch.randn(1, 64, f, h, w, generator=torch.Generator().manual_seed(t + f)).to(dev)
    if f == 1:
        qk = qk[:, :, 0].contiguous()
    bank = ev.MemoryBank(1, 64, 512, h, w, t, dev, keep_reference_layout=False)
    bank.write_frames(0, mk.to(dev), mv.to(dev))
    row = []
    for r in (1, 2, 3):
        lib.evavos_stage_timing(1)
        acc, reps = np.zeros(4), 10
        for i in range(reps + 3):
            ev.memory_read(bank, qk, 50, sample_stride=r)
            ms = (ctypes.c_float * 4)()
            lib.evavos_stage_timing_read(ms)
            if i >= 3:
                acc += np.array(list(ms))
        lib.evavos_stage_timing(0)
        us = acc / reps * 1e3
        row.append(f"R={r}: {us[0]:7.1f} + {us[2]:6.1f}")
    print(f"[{t * h * w:7d} positions x {f * h * w:5d} queries] filter + finalize us  " + "   ".join(row), flush=True)
    del bank
