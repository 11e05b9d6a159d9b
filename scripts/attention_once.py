"""One attention read at 68 x 120 per form (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402

h, w, n_vec = 68, 120, int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = torch.Generator().manual_seed(3)
mk, qk = torch.randn(1, 64, 1, h, w, generator=g).cuda(), torch.randn(1, 64, h, w, generator=g).cuda()
vec = torch.rand(n_vec, h * w, generator=g).cuda()
for form in ("simt", "tensor"):
    os.environ["EVAVOS_ATTENTION_PATH"] = form
    for _ in range(2):
        ev.attention_readout(mk, qk, vec)
torch.cuda.synchronize()
