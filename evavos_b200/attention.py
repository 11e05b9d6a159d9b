"""Attention read of the fusion path on the sm_100a kernels.

``attention_readout`` replaces ``AttentionMemory.forward`` followed by the two ``(b,1,HW) @ W`` products of
``PropagationNetwork.get_attention`` (mivos/model/propagation/prop_net.py:117-138, 198-207): the (HW, HW) softmax
matrix is never materialised.
"""
from __future__ import annotations

import torch

from . import _lib
from .memory_reader import _require_cuda, _workspace


def attention_readout(mem_key: torch.Tensor, query_key: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    """out[c, q] = sum_n vec[c, n] * softmax_n(affinity(mem_key_n, query_key_q)).

    mem_key (1,CK,1,H,W) or (1,CK,H,W): the ONE memory frame; query_key (1,CK,H,W); vec (C, H*W) fp32 rows.
    Returns (C, H*W) fp32.
    """
    lib = _lib.load()
    _require_cuda(mem_key, "memory key")
    _require_cuda(query_key, "query key")
    if mem_key.shape[0] != 1 or query_key.shape[0] != 1:
        raise ValueError("attention_readout: batch must be 1 (as in the reference's get_attention)")
    ck = mem_key.shape[1]
    mk = mem_key.to(torch.float32).reshape(ck, -1)
    qk = query_key.to(torch.float32).reshape(ck, -1)
    vec = vec.to(torch.float32)
    if mk.stride(1) != 1:
        mk = mk.contiguous()
    if qk.stride(1) != 1:
        qk = qk.contiguous()
    if vec.dim() != 2 or vec.shape[1] != mk.shape[1]:
        raise ValueError(f"vec {tuple(vec.shape)} does not match the {mk.shape[1]} memory positions")
    if vec.stride(1) != 1:
        vec = vec.contiguous()
    dev = mk.device
    n_vec, n_mem, n_query = vec.shape[0], mk.shape[1], qk.shape[1]
    out = torch.empty((n_vec, n_query), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        # the kernel carries at most 32 mask rows per query: more objects (> 15) take several launches
        for r0 in range(0, n_vec, 32):
            rows = min(32, n_vec - r0)
            need = lib.evavos_attention_workspace_bytes(rows, n_mem, n_query, n_sm)
            ws = _workspace.get(dev, max(int(need), 1))
            _lib.check(lib.evavos_attention_readout(mk.data_ptr(), mk.stride(0), qk.data_ptr(), qk.stride(0),
                                                    vec[r0:].data_ptr(), vec.stride(0), rows, ck, n_mem, n_query,
                                                    out[r0:].data_ptr(), out.stride(0), ws.data_ptr(), ws.numel(), n_sm,
                                                    _lib.current_stream_ptr(dev)))
    return out
