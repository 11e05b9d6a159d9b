"""Host-buffer form of the read: the call a user without device tensors makes.

Wraps ``evavos_memread_host`` (include/evavos.h): inputs are HOST tensors in the reference
layouts, the library copies them to the device, builds the shadow, reads and copies back.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def memory_read_host(mem_key: torch.Tensor, query_key: torch.Tensor, mem_value: torch.Tensor, top_k: int = 50,
                     out: torch.Tensor | None = None, path: int = _lib.PATH_AUTO, want_topk: bool = False):
    """mem_key (CK,N) | (1,CK,T,H,W), query_key (CK,HW) | (1,CK,H,W), mem_value (K,CV,N) | (K,CV,T,H,W): CPU fp32.

    Returns (h2d_bytes, d2h_bytes) and fills ``out`` (K,CV,HW) (allocated when None and returned
    as a third element).  Pinned tensors make the copies asynchronous inside the call.
    """
    lib = _lib.load()
    for t_, name in ((mem_key, "mem_key"), (query_key, "query_key"), (mem_value, "mem_value")):
        if t_.is_cuda or t_.dtype != torch.float32:
            raise ValueError(f"{name} must be a CPU float32 tensor (use memory_read for device tensors)")
    ck = mem_key.shape[1] if mem_key.dim() == 5 else mem_key.shape[0]
    mk = mem_key.reshape(ck, -1).contiguous()
    qk = query_key.reshape(ck, -1).contiguous()
    k, cv = mem_value.shape[0], mem_value.shape[1]
    mv = mem_value.reshape(k, cv, -1).contiguous()
    n_pos, nq = mk.shape[1], qk.shape[1]
    made = out is None
    if made:
        out = torch.empty((k, cv, nq), dtype=torch.float32)
    if not out.is_contiguous() or out.numel() != k * cv * nq:
        raise ValueError("out must be a contiguous (K,CV,HW) float32 CPU tensor")
    idx = torch.empty((nq, top_k), dtype=torch.int32) if want_topk else None
    wgt = torch.empty((nq, top_k), dtype=torch.float32) if want_topk else None
    h2d, d2h = ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check(lib.evavos_memread_host(mk.data_ptr(), qk.data_ptr(), mv.data_ptr(), k, ck, cv, n_pos, nq, int(top_k),
                                       int(path), out.data_ptr(), idx.data_ptr() if want_topk else None,
                                       wgt.data_ptr() if want_topk else None, ctypes.byref(h2d), ctypes.byref(d2h)))
    res = (h2d.value, d2h.value)
    if made:
        res = res + (out,)
    if want_topk:
        res = res + (idx, wgt)
    return res
