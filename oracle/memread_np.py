"""numpy restatement of the reference memory read, used as the parity checker.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the
reference lines it follows (paths relative to /root/reference).

The checker works in float64 so that it can tell a genuine top-k difference
from a near-tie that the reference's own fp32 rounding could flip either way.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


# --------------------------------------------------------------------------- #
# affinity + top-k softmax                                                     #
# --------------------------------------------------------------------------- #
def affinity_scores(mk: np.ndarray, qk: np.ndarray, dtype=np.float64) -> np.ndarray:
    """S[n, q] = (-|mk_n|^2 + 2 mk_n.qk_q - |qk_q|^2) / sqrt(CK).

    Follows mivos/model/propagation/prop_net.py:80-90 (EvalMemoryReader.get_affinity).
    mk: (CK, N)  memory keys flattened over (T, H, W);  qk: (CK, HW).
    Returns (N, HW).
    """
    mk = np.asarray(mk, dtype=dtype)
    qk = np.asarray(qk, dtype=dtype)
    ck = mk.shape[0]
    a = (mk * mk).sum(0)[:, None]          # prop_net.py:86
    b = 2.0 * (mk.T @ qk)                  # prop_net.py:87
    c = (qk * qk).sum(0)[None, :]          # prop_net.py:88
    return (-a + b - c) / math.sqrt(ck)    # prop_net.py:90


@dataclass
class TopK:
    idx: np.ndarray      # (HW, k) int64, memory positions, best first
    score: np.ndarray    # (HW, k) float64 affinity of those positions
    weight: np.ndarray   # (HW, k) float64 softmax weights over the k survivors
    kth_gap: np.ndarray  # (HW,)   score[k-1] - score of the (k+1)-th best (tie detector)


def topk_softmax(scores: np.ndarray, k: int) -> TopK:
    """Top-k over the memory axis, softmax over the survivors.

    Follows softmax_w_g_top (prop_net.py:46-72, the ``gauss is None`` branch):
    topk(x, k, dim=1) sorted, exp(values - values[:,0]), normalise.
    Ties are ordered by ascending memory index (torch leaves them unspecified).
    scores: (N, HW).
    """
    n, hw = scores.shape
    if k > n:
        # torch.topk raises "selected index k out of range" (prop_net.py:53)
        raise RuntimeError("selected index k out of range")
    order = np.lexsort((np.arange(n)[:, None].repeat(hw, 1), -scores), axis=0)  # by -score then idx
    top = order[:k].T                                   # (HW, k)
    sc = np.take_along_axis(scores.T, top, axis=1)      # (HW, k)
    e = np.exp(sc - sc[:, :1])                          # prop_net.py:54
    w = e / e.sum(1, keepdims=True)                     # prop_net.py:56-57
    if n > k:
        nxt = np.take_along_axis(scores.T, order[k:k + 1].T, axis=1)[:, 0]
        gap = sc[:, -1] - nxt
    else:
        gap = np.full(hw, np.inf)
    return TopK(top.astype(np.int64), sc, w, gap)


def dense_affinity(idx: np.ndarray, weight: np.ndarray, n: int) -> np.ndarray:
    """x.zero_().scatter_(1, indices, x_exp)  (prop_net.py:60) -> (N, HW)."""
    hw, k = idx.shape
    out = np.zeros((n, hw), dtype=weight.dtype)
    out[idx, np.arange(hw)[:, None].repeat(k, 1)] = weight
    return out


def readout(idx: np.ndarray, weight: np.ndarray, mv: np.ndarray) -> np.ndarray:
    """mem = mv.view(CV, N) @ affinity  (prop_net.py:108-115), sparse form.

    mv: (K, CV, N) or (CV, N); returns (K, CV, HW) / (CV, HW) in float64.
    """
    mv = np.asarray(mv, dtype=np.float64)
    squeeze = mv.ndim == 2
    if squeeze:
        mv = mv[None]
    g = mv[:, :, idx]                                   # (K, CV, HW, k)
    out = (g * np.asarray(weight, np.float64)[None, None]).sum(-1)
    return out[0] if squeeze else out


def memory_read(mk, qk, mv, k=50):
    """get_affinity + per-object readout (prop_net.py:179-187)."""
    s = affinity_scores(mk, qk)
    tk = topk_softmax(s, k)
    return tk, readout(tk.idx, tk.weight, mv)


# --------------------------------------------------------------------------- #
# tie-aware comparison                                                         #
# --------------------------------------------------------------------------- #
def compare_topk(test_idx: np.ndarray, scores64: np.ndarray, k: int, tie_tol: float):
    """Compare a candidate top-k index set against the fp64 scores.

    A query passes when every selected position has a score >= (k-th best score
    - tie_tol) and every position with score > (k-th best + tie_tol) is selected,
    i.e. the two sets may differ only among positions whose scores are within
    ``tie_tol`` of the k-th best ("ties" at the working precision).
    Returns (n_exact_equal_sets, n_tie_only_diffs, n_bad, bad_query_list).
    """
    n, hw = scores64.shape
    part = -np.partition(-scores64, k - 1, axis=0)[k - 1]       # k-th best per query
    exact = tie = bad = 0
    bad_q = []
    ref_sets = np.argsort(-scores64, axis=0, kind="stable")[:k].T
    for q in range(hw):
        t = np.asarray(test_idx[q], dtype=np.int64)
        if len(set(t.tolist())) != k or t.min() < 0 or t.max() >= n:
            bad += 1
            bad_q.append(q)
            continue
        if set(t.tolist()) == set(ref_sets[q].tolist()):
            exact += 1
            continue
        col = scores64[:, q]
        ok_low = (col[t] >= part[q] - tie_tol).all()
        must = np.nonzero(col > part[q] + tie_tol)[0]
        ok_must = np.isin(must, t).all()
        if ok_low and ok_must:
            tie += 1
        else:
            bad += 1
            bad_q.append(q)
    return exact, tie, bad, bad_q


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# --------------------------------------------------------------------------- #
# soft aggregation                                                             #
# --------------------------------------------------------------------------- #
def aggregate_wbg(prob: np.ndarray, keep_bg: bool = False, hard: bool = False,
                  dtype=np.float64) -> np.ndarray:
    """Follows mivos/model/aggregate.py:22-37.

    prob: (K, 1, h, w) object probabilities.  Returns (K+1, 1, h, w) if keep_bg
    else (K, 1, h, w).  The clamp constants are the float32 values the reference
    uses (1e-7 and 1-1e-7 rounded to fp32) so fp64 evaluation matches its limits.
    """
    p = np.asarray(prob, dtype=dtype)
    bg = np.prod(1.0 - p, axis=0, keepdims=True)            # aggregate.py:25
    allp = np.concatenate([bg, p], 0)
    lo = dtype(np.float32(1e-7))
    hi = dtype(np.float32(1 - 1e-7))
    allp = np.clip(allp, lo, hi)                            # aggregate.py:27
    logits = np.log(allp / (1.0 - allp))                    # aggregate.py:28
    if hard:
        logits = logits * 1000.0                            # aggregate.py:30-32
    logits = logits - logits.max(0, keepdims=True)
    e = np.exp(logits)
    sm = e / e.sum(0, keepdims=True)                        # aggregate.py:34-37
    return sm if keep_bg else sm[1:]


# --------------------------------------------------------------------------- #
# memory bank + padding                                                        #
# --------------------------------------------------------------------------- #
def bank_append(keys: np.ndarray, values: np.ndarray, slot: int,
                key_frame: np.ndarray, value_frame: np.ndarray) -> None:
    """keys[:,:,slot] = k16 ; values[:,:,slot] = v  (inference_core.py:174-177).

    keys (1,CK,T,H,W), values (K,CV,T,H,W); key_frame (1,CK,H,W) or (1,CK,1,H,W).
    """
    keys[:, :, slot] = np.asarray(key_frame).reshape(keys.shape[0], keys.shape[1], *keys.shape[3:])
    values[:, :, slot] = np.asarray(value_frame).reshape(values.shape[0], values.shape[1], *values.shape[3:])


def pad_amounts(h: int, w: int, d: int = 16):
    """(lw, uw, lh, uh) of pad_divide_by (mivos/tensor_util.py:62-80)."""
    new_h = h + d - h % d if h % d > 0 else h
    new_w = w + d - w % d if w % d > 0 else w
    lh, uh = int((new_h - h) / 2), int(new_h - h) - int((new_h - h) / 2)
    lw, uw = int((new_w - w) / 2), int(new_w - w) - int((new_w - w) / 2)
    return (int(lw), int(uw), int(lh), int(uh))


# --------------------------------------------------------------------------- #
# memory-axis sharded read (the multi-GPU exchange, restated on one host)      #
# --------------------------------------------------------------------------- #
def sharded_memory_read(mk, qk, mv, k, owners):
    """Reference semantics of the THW-sharded read.

    ``owners``: list of index arrays, one per shard, partitioning range(N).
    Each shard computes its local top-k (score, global idx); the union is merged
    to the global top-k, softmax weights use the global max and denominator, and
    every shard contributes the partial readout of the winners it owns.
    Returns (TopK of the merged result, summed readout) - must equal memory_read.
    """
    mk = np.asarray(mk, np.float64)
    s = affinity_scores(mk, qk)
    n, hw = s.shape
    cand_idx, cand_sc = [], []
    for own in owners:
        kk = min(k, len(own))
        loc = topk_softmax(s[own], kk)
        cand_idx.append(np.asarray(own)[loc.idx])
        cand_sc.append(loc.score)
    ci = np.concatenate(cand_idx, 1)
    cs = np.concatenate(cand_sc, 1)
    order = np.lexsort((ci, -cs), axis=1)[:, :k]
    gi = np.take_along_axis(ci, order, 1)
    gs = np.take_along_axis(cs, order, 1)
    e = np.exp(gs - gs[:, :1])
    w = e / e.sum(1, keepdims=True)
    mv = np.asarray(mv, np.float64)
    total = np.zeros((mv.shape[0], mv.shape[1], hw))
    for own in owners:
        mask = np.isin(gi, own)
        total += readout(gi, np.where(mask, w, 0.0), mv)
    return TopK(gi, gs, w, np.zeros(hw)), total


def attention_readout(mk: np.ndarray, qk: np.ndarray, vec: np.ndarray) -> np.ndarray:
    """vec @ softmax_n(S), fp64: AttentionMemory.forward (prop_net.py:123-138) followed by the
    vector-matrix products of get_attention (:204-205).

    mk (CK, M), qk (CK, Q), vec (C, M) -> (C, Q).
    """
    s = affinity_scores(mk, qk)                       # (M, Q) fp64, (-a + b - c) / sqrt(CK)
    s = s - s.max(axis=0, keepdims=True)
    w = np.exp(s)
    w /= w.sum(axis=0, keepdims=True)
    return vec.astype(np.float64) @ w
