"""torch-CPU fp32 port of the reference's dense memory read (TEST INFRASTRUCTURE ONLY).

This is the op-for-op CPU port that ``bench.py`` times as ``cpu_baseline`` /
``--impl reference`` on the GPU box (the reference itself is Python and lives in
/root/reference, which does not exist there).  It executes the same ATen op
sequence as the reference on the same layouts: a dense (N, HW) fp32 affinity,
``topk`` along the memory axis, in-place zero + scatter, and one dense ``bmm``
per object.  It is validated against the real reference in
``oracle/make_golden.py`` (bit-identical outputs on CPU) and against the
committed golden vectors in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math

import torch


def dense_topk_affinity(mem_key: torch.Tensor, query_key: torch.Tensor, top_k: int = 50) -> torch.Tensor:
    """Dense top-k-softmax affinity, (1, N, HW) fp32.

    mem_key (1,CK,T,H,W), query_key (1,CK,H,W).  Op sequence of
    EvalMemoryReader.get_affinity + softmax_w_g_top with gauss=None
    (mivos/model/propagation/prop_net.py:80-90 and :46-62).
    """
    ck = mem_key.shape[1]
    m = mem_key.flatten(start_dim=2)                    # (1, CK, N)   view, prop_net.py:83
    q = query_key.flatten(start_dim=2)                  # (1, CK, HW)
    m_sq = m.pow(2).sum(1).unsqueeze(2)                 # (1, N, 1)    :86
    cross = 2 * (m.transpose(1, 2) @ q)                 # (1, N, HW)   :87
    q_sq = q.pow(2).sum(1).unsqueeze(1)                 # (1, 1, HW)   :88
    aff = (-m_sq + cross - q_sq) / math.sqrt(ck)        # :90
    vals, pos = torch.topk(aff, k=top_k, dim=1)         # :53
    e = torch.exp(vals - vals[:, 0])                    # :54
    e /= torch.sum(e, dim=1, keepdim=True)              # :56-57
    aff.zero_().scatter_(1, pos, e.type(aff.dtype))     # :60
    return aff


def dense_readout(affinity: torch.Tensor, mem_value: torch.Tensor) -> torch.Tensor:
    """(1,CV,T,H,W) x (1,N,HW) -> (1,CV,H,W); prop_net.py:108-115 (bmm over a strided view)."""
    b, cv, t, h, w = mem_value.shape
    out = torch.bmm(mem_value.view(b, cv, t * h * w), affinity)
    # (bench.py times a slice of the query columns when the full frame is too slow: keep it flat then)
    return out.view(b, cv, h, w) if out.shape[-1] == h * w else out


def memory_read(mem_key, query_key, mem_value, top_k: int = 50) -> torch.Tensor:
    """Affinity once, readout per object (prop_net.py:180-187). Returns (K,CV,H,W)."""
    aff = dense_topk_affinity(mem_key, query_key, top_k)
    per_obj = [dense_readout(aff, mem_value[i:i + 1]) for i in range(mem_value.shape[0])]
    return torch.cat(per_obj, 0)


def aggregate_wbg(prob: torch.Tensor, keep_bg: bool = False, hard: bool = False) -> torch.Tensor:
    """Soft aggregation with a product-of-complements background (aggregate.py:22-37)."""
    bg = torch.prod(1 - prob, dim=0, keepdim=True)
    stacked = torch.cat([bg, prob], 0).clamp(1e-7, 1 - 1e-7)
    logit = torch.log(stacked / (1 - stacked))
    if hard:
        logit *= 1000
    sm = torch.softmax(logit, dim=0)
    return sm if keep_bg else sm[1:]
