"""The decoder-tail kernels at the sizes of a 5-frame 480p segment (for ncu captures): up_8_4's tails, fp32 and bf16."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evavos_b200.decoder_ops import bias_residual_, upsample2x_add_  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
for dt in (torch.float32, torch.bfloat16):
    y = torch.randn(5, 256, 120, 216, device="cuda", generator=g).to(dt).contiguous(memory_format=torch.channels_last)
    r = torch.randn(5, 256, 120, 216, device="cuda", generator=g).to(dt).contiguous(memory_format=torch.channels_last)
    x = torch.randn(5, 256, 60, 108, device="cuda", generator=g).to(dt).contiguous(memory_format=torch.channels_last)
    b = torch.randn(256, device="cuda", generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        flush.zero_()                       # y / r / x out of L2 before each kernel
        bias_residual_(y, b, r, relu=True)
        flush.zero_()
        upsample2x_add_(y, b, x)
torch.cuda.synchronize()
