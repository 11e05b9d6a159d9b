"""Where the end-to-end interact() of cfg3 spends its time: wall clock vs summed kernel time vs launch count.

python scripts/profile_cfg3.py [amp] [bench]   (amp: bf16 autocast; bench: cudnn.benchmark = True)
Prints the wall time of one 32-frame video, the number of device kernels, their summed duration and the top
kernels by total time (torch.profiler, one video after two warm ones).  Not a bench value.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from evavos_b200.networks import seeded_init  # noqa: E402


def main():
    amp = "amp" in sys.argv
    torch.backends.cudnn.benchmark = "bench" in sys.argv
    dev = torch.device("cuda", 0)
    torch.set_grad_enabled(False)
    prop, fuse = ev.PropagationNetwork().eval().to(dev), ev.FusionNet().eval().to(dev)
    seeded_init(prop, 1001)
    seeded_init(fuse, 1002)
    t, h, w, k = 32, 480, 854, 1
    g = torch.Generator().manual_seed(7)
    video = torch.rand(1, t, 3, h, w, generator=g)
    mask = (torch.rand(1, 1, h // 8, (w + 7) // 8, generator=g) > 0.6).float()
    mask = mask.repeat_interleave(8, 2).repeat_interleave(8, 3)[:, :, :h, :w]
    kw = dict(amp=amp, fold_bn="nofold" not in sys.argv)
    if "cl" in sys.argv:
        kw["channels_last"] = True
    if "graphs" in sys.argv:
        kw["cuda_graphs"] = True
    if "tf32off" in sys.argv:
        torch.backends.cudnn.allow_tf32 = False

    def one():
        proc = ev.InferenceCore(prop, fuse, video, k, device=dev, **kw)
        return proc.interact(mask, 0)

    for _ in range(2):
        one()
    torch.cuda.synchronize()
    walls = []
    for _ in range(3):
        c0 = time.perf_counter()
        one()
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - c0)
    print(f"{' '.join(sys.argv[1:])}: wall per video "
          f"{1e3 * min(walls):.1f} ms -> {(t - 1) / min(walls):.0f} frames/s")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        one()
        torch.cuda.synchronize()
    ev_list = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = sum(e.device_time for e in ev_list)
    print(f"device events {len(ev_list)}, summed device time {tot / 1e3:.1f} ms")
    agg = {}
    for e in ev_list:
        a = agg.setdefault(e.name[:200], [0, 0.0])
        a[0] += 1
        a[1] += e.device_time
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f"{us / 1e3:8.2f} ms {n:6d} x  {name[:60]} ... {name[-100:] if len(name) > 60 else ''}")


if __name__ == "__main__":
    main()
