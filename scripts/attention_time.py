"""Both forms of the attention read (csrc/attention.cu: CUDA cores / 3 x TF32 mma.sync) timed with CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402

g = torch.Generator().manual_seed(3)
for h, w in ((30, 54), (40, 72), (48, 90), (68, 120)):
    for n_vec in (4, 8, 12):
        mk, qk = torch.randn(1, 64, 1, h, w, generator=g).cuda(), torch.randn(1, 64, h, w, generator=g).cuda()
        vec = torch.rand(n_vec, h * w, generator=g).cuda()
        row = []
        for form in ("simt", "tensor"):
            os.environ["EVAVOS_ATTENTION_PATH"] = form
            for _ in range(3):
                out = ev.attention_readout(mk, qk, vec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                out = ev.attention_readout(mk, qk, vec)
            e1.record()
            torch.cuda.synchronize()
            row.append((e0.elapsed_time(e1) * 50, out))
        diff = (row[0][1] - row[1][1]).abs().max().item()
        gf = 2 * (h * w) ** 2 * 64 / 1e9
        print(f"{h}x{w} n_vec {n_vec:2d}: simt {row[0][0]:8.1f} us  tensor {row[1][0]:8.1f} us  ({gf / row[1][0] * 1e3:6.1f} TF/s useful)"
              f"  max |simt - tensor| {diff:.2e}", flush=True)
