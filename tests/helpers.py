"""Shared helpers for the parity tests (oracle = checker only)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Two positions count as tied when their fp64 affinities differ by less than this.  The reference
# computes the affinity in fp32 with an SGEMM over 64 channels of magnitude ~|k||q|/8, so its own
# ranking is only defined up to a few 1e-6 relative to scores of order 10-30.
TIE_TOL = 5e-5


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def synth(seed, ck, cv, t, h, w, k, scale=1.0):
    """SURVEY.md 8d synthetic inputs (identical to oracle/make_golden.py:synth)."""
    g = torch.Generator().manual_seed(seed)
    mk = torch.randn(1, ck, t, h, w, generator=g) * scale
    qk = torch.randn(1, ck, h, w, generator=g) * scale
    mv = torch.randn(k, cv, t, h, w, generator=g)
    return mk, qk, mv
