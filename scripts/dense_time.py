"""Time of a read whose every candidate list overflows (near-constant keys): filter + finalize + exact tiled pass."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from evavos_b200 import _lib  # noqa: E402
from evavos_b200.memory_reader import last_overflow_count  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
h, w = 30, 54
for t in (3, 5, 8):
    base = torch.randn(1, 64, 1, 1, 1, generator=g, device=dev)
    bank = ev.MemoryBank(1, 64, 512, h, w, t, dev, keep_reference_layout=False)
    bank.write_frames(0, base + 1e-3 * torch.randn(1, 64, t, h, w, generator=g, device=dev),
                      torch.randn(1, 512, t, h, w, generator=g, device=dev))
    qk = 0.9 * base + 1e-3 * torch.randn(1, 64, 5, h, w, generator=g, device=dev)
    res = []
    for path in (_lib.PATH_TENSOR_DENSE, _lib.PATH_SIMT):
        for _ in range(3):
            ev.memory_read(bank, qk, 50, path=path, want_readout=False, want_topk=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ev.memory_read(bank, qk, 50, path=path, want_readout=False, want_topk=True)
        e1.record()
        torch.cuda.synchronize()
        res.append(1e3 * e0.elapsed_time(e1) / 10)
        if path == _lib.PATH_TENSOR_DENSE:
            n_over = last_overflow_count()
    print(f"{t * h * w} positions x 8100 queries, {n_over} overflowed: tensor path + exact tiled pass {res[0]:.0f} us; "
          f"exact SIMT selection alone {res[1]:.0f} us", flush=True)
