"""Golden vectors that need the reference's networks (build container only; see make_golden.py).

1. tests/golden/propnet_state_dict.json - parameter/buffer names and shapes of the reference
   PropagationNetwork and FusionNet (stcn.pth / fusion.pth must load into ours unchanged).
2. tests/golden/e2e_*.npz - InferenceCore.interact round trips with seeded random weights on a
   synthetic video: the reference's `prob` tensors and output masks.  The seeded weights are produced by
   evavos_b200.networks.seeded_init, applied to BOTH the reference networks and ours through the shared
   state dict, so the GPU test rebuilds the same weights without the reference.
"""
from __future__ import annotations

import json
import zlib
import os
import sys
from unittest import mock

import numpy as np
import torch
import torchvision

REF = os.environ.get("EVAVOS_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def build_reference_nets():
    from mivos.model.propagation import mod_resnet, modules
    from mivos.model.propagation.prop_net import PropagationNetwork
    from mivos.model.fusion_net import FusionNet
    # the constructors download ImageNet weights; offline we only need the architecture
    real_resnet50 = torchvision.models.resnet50
    with mock.patch.object(mod_resnet.model_zoo, "load_url", return_value={}), \
            mock.patch.object(modules.models, "resnet50", lambda weights=None: real_resnet50(weights=None)):
        prop = PropagationNetwork()
    return prop.eval(), FusionNet().eval()


def main():
    torch.set_grad_enabled(False)
    prop, fuse = build_reference_nets()
    spec = {"prop": {k: list(v.shape) for k, v in prop.state_dict().items()},
            "fuse": {k: list(v.shape) for k, v in fuse.state_dict().items()}}
    with open(os.path.join(OUT, "propnet_state_dict.json"), "w") as f:
        json.dump(spec, f, indent=0)
    print("state dict entries:", len(spec["prop"]), len(spec["fuse"]))
    if "--spec-only" in sys.argv:
        return

    from evavos_b200.networks import seeded_init
    from mivos.inference_core import InferenceCore
    gain = float(os.environ.get("EVAVOS_INIT_GAIN", "0.6"))
    seeded_init(prop, 1001, gain)
    seeded_init(fuse, 1002, gain)

    cases = [
        # tag, T, h, w, K, mem_freq, interactions [(frame, scribble)]
        ("e2e_k1", 7, 120, 150, 1, 2, [(0, False), (5, False)]),
        ("e2e_k2", 6, 100, 140, 2, 3, [(2, True)]),
    ]
    for tag, t, h, w, k, mem_freq, inter in cases:
        g = torch.Generator().manual_seed(zlib.crc32(tag.encode()) % 1000 + 7)
        images = torch.rand(1, t, 3, h, w, generator=g)
        proc = InferenceCore(prop, fuse, images, k, mem_freq=mem_freq, device="cpu")
        out = {"images": images.numpy(), "num_objects": k, "mem_freq": mem_freq}
        for n, (frame, scribble) in enumerate(inter):
            if scribble:
                lab = torch.randint(0, k + 1, ((h + 3) // 4, (w + 3) // 4), generator=g)
                lab = lab.repeat_interleave(4, 0).repeat_interleave(4, 1)[:h, :w]
                mask = torch.stack([(lab == c).float() for c in range(k + 1)], 0).unsqueeze(1)   # (k+1,1,h,w) with bg
            else:
                blob = (torch.rand(1, 1, (h + 7) // 8, (w + 7) // 8, generator=g) > 0.6).float()
                mask = blob.repeat_interleave(8, 2).repeat_interleave(8, 3)[:, :, :h, :w]        # (1,1,h,w) object only
            np_masks = proc.interact(mask, frame, scribble=scribble)
            out[f"mask_{n}"] = mask.numpy()
            out[f"frame_{n}"] = frame
            out[f"scribble_{n}"] = int(scribble)
            out[f"np_masks_{n}"] = np_masks
            out[f"prob_{n}"] = proc.prob.numpy().copy()
        out["n_interactions"] = len(inter)
        out["pad"] = np.array(proc.pad)
        np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
        print(tag, "written; prob range", float(proc.prob.min()), float(proc.prob.max()))


if __name__ == "__main__":
    main()
