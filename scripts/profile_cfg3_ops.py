"""Which torch ops (with input shapes) the elementwise / copy kernels of a cfg3 video come from (amp + graphs off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from evavos_b200.networks import seeded_init  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = torch.device("cuda", 0)
torch.set_grad_enabled(False)
prop, fuse = ev.PropagationNetwork().eval().to(dev), ev.FusionNet().eval().to(dev)
seeded_init(prop, 1001)
seeded_init(fuse, 1002)
t, h, w = 32, 480, 854
g = torch.Generator().manual_seed(7)
video = torch.rand(1, t, 3, h, w, generator=g)
mask = (torch.rand(1, 1, h // 8, (w + 7) // 8, generator=g) > 0.6).float().repeat_interleave(8, 2).repeat_interleave(8, 3)[:, :, :h, :w]
amp = "amp" in sys.argv


def one():
    return ev.InferenceCore(prop, fuse, video, 1, device=dev, amp=amp, channels_last=True).interact(mask, 0)


for _ in range(2):
    one()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    one()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.key in ("aten::add", "aten::add_", "aten::copy_", "aten::_to_copy", "aten::contiguous", "aten::clone", "aten::cat",
                 "aten::relu", "aten::clamp_min", "aten::mul", "aten::sigmoid", "aten::upsample_bilinear2d",
                 "aten::repeat_interleave", "aten::index_put_", "aten::slice", "aten::to"):
        rows.append((e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:110]))
for us, n, key, shapes in sorted(rows, reverse=True)[:32]:
    print(f"{us / 1e3:7.2f} ms {n:5d} x {key:24s} {shapes}")
