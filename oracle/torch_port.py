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
import torch.nn.functional as F


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


def attention_weights(mem_key: torch.Tensor, query_key: torch.Tensor) -> torch.Tensor:
    """AttentionMemory.forward, prop_net.py:123-138: full softmax over the ONE memory frame.

    mem_key (B,CK,1,H,W) (or any (B,CK,...) that flattens to THW), query_key (B,CK,H,W) -> W (B,THW,HW).
    """
    ck = mem_key.shape[1]
    mk = mem_key.flatten(start_dim=2)
    qk = query_key.flatten(start_dim=2)
    a = mk.pow(2).sum(1).unsqueeze(2)
    b = 2 * (mk.transpose(1, 2) @ qk)
    c = qk.pow(2).sum(1).unsqueeze(1)
    affinity = (-a + b - c) / math.sqrt(ck)
    return F.softmax(affinity, dim=1)


def attention_lowres(mem_key, pos_mask, neg_mask, query_key) -> torch.Tensor:
    """The stride-16 part of get_attention, prop_net.py:198-208: (b,2,nh,nw) positive / negative attention."""
    b, _, h, w = pos_mask.shape
    nh, nw = h // 16, w // 16
    W = attention_weights(mem_key, query_key)
    pos_map = F.interpolate(pos_mask, size=(nh, nw), mode="area").view(b, 1, nh * nw) @ W
    neg_map = F.interpolate(neg_mask, size=(nh, nw), mode="area").view(b, 1, nh * nw) @ W
    return torch.cat([pos_map, neg_map], 1).reshape(b, 2, nh, nw)


def get_attention(mem_key, pos_mask, neg_mask, query_key) -> torch.Tensor:
    """PropagationNetwork.get_attention, prop_net.py:198-211: (b,2,h,w)."""
    h, w = pos_mask.shape[-2:]
    return F.interpolate(attention_lowres(mem_key, pos_mask, neg_mask, query_key), mode="bilinear", size=(h, w),
                         align_corners=False)
