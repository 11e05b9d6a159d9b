"""Generate tests/golden/attention.npz by running the UNMODIFIED reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_attention.py

Covers the fusion path's attention read: AttentionMemory.forward (prop_net.py:117-138) and
PropagationNetwork.get_attention (prop_net.py:198-211).  get_attention only touches ``self.get_W``, so it is
called unbound on a stub that holds the reference's own AttentionMemory (building the whole network would need the
ImageNet download of its constructor).  While generating, the script asserts that oracle/torch_port.py reproduces
the reference bit-for-bit and that oracle/memread_np.py (fp64) agrees within fp32 rounding.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("EVAVOS_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mivos.model.propagation.prop_net import AttentionMemory, PropagationNetwork  # noqa: E402  (reference)

from oracle import memread_np as onp  # noqa: E402
from oracle import torch_port as port  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def blobs(g, b, h, w):
    """Sparse positive / negative interaction differences in [0, 1] (inference_core.py:212-214)."""
    m = torch.zeros(b, 1, h, w)
    for i in range(b):
        for _ in range(2):
            y, x = int(torch.randint(0, h - 20, (1,), generator=g)), int(torch.randint(0, w - 20, (1,), generator=g))
            m[i, 0, y:y + 20, x:x + 20] = torch.rand(20, 20, generator=g)
    return m


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    stub = types.SimpleNamespace()
    stub.attn_memory = AttentionMemory(50)
    stub.get_W = lambda mk16, qk16: stub.attn_memory(mk16, qk16)
    out = {}
    cases = (("b2", 2, 96, 144, 1.0, 11), ("b4_peaky", 4, 112, 160, 2.0, 12), ("b1_flat", 1, 96, 144, 0.1, 13))
    for name, b, h, w, scale, seed in cases:
        g = torch.Generator().manual_seed(seed)
        nh, nw = h // 16, w // 16
        mk = torch.randn(1, 64, 1, nh, nw, generator=g) * scale
        qk = torch.randn(1, 64, nh, nw, generator=g) * scale
        pos, neg = blobs(g, b, h, w), blobs(g, b, h, w)
        ref = PropagationNetwork.get_attention(stub, mk, pos, neg, qk)            # (b,2,h,w), the reference itself
        W = stub.attn_memory(mk, qk)
        assert torch.equal(port.attention_weights(mk, qk), W)
        assert torch.equal(port.get_attention(mk, pos, neg, qk), ref)
        low = port.attention_lowres(mk, pos, neg, qk)
        vec = torch.stack([torch.nn.functional.interpolate(pos, size=(nh, nw), mode="area").view(b, -1),
                           torch.nn.functional.interpolate(neg, size=(nh, nw), mode="area").view(b, -1)], 1)
        o64 = onp.attention_readout(mk.reshape(64, -1).numpy(), qk.reshape(64, -1).numpy(),
                                    vec.reshape(2 * b, -1).numpy()).reshape(b, 2, nh, nw)
        err = np.abs(o64 - low.numpy()).max()
        assert err < 2e-6, err
        print(f"  {name}: port == reference, fp64 oracle within {err:.2e}")
        out.update({f"{name}_mk": mk.numpy(), f"{name}_qk": qk.numpy(), f"{name}_pos": pos.numpy(),
                    f"{name}_neg": neg.numpy(), f"{name}_lowres": low.numpy(), f"{name}_attn": ref.numpy()})
    np.savez_compressed(os.path.join(OUT, "attention.npz"), **out)
    print("attention golden written to", OUT)


if __name__ == "__main__":
    main()
