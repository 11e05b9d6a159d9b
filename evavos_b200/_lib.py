"""ctypes binding of libevavos_sm100.so (the C ABI declared in include/evavos.h).

There is deliberately no fallback: if the shared library is missing or its ABI does not
match, importing a kernel entry point raises.  Build it with ``python -m evavos_b200.build``
(or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
import threading

ABI_VERSION = 2
F32, BF16 = 0, 1
PATH_AUTO, PATH_TENSOR, PATH_SIMT, PATH_TENSOR_DENSE = 0, 1, 2, 3
TILE_POS = 128
TILE_BYTES = 20480
MAX_TOPK = 128
ERR_TOPK_RANGE = -5

LIB_NAME = "libevavos_sm100.so"
LIB_PATH = os.environ.get("EVAVOS_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

_c_i32, _c_i64, _c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p


class BankShadow(ctypes.Structure):
    """struct EvavosBankShadow (include/evavos.h)."""
    _fields_ = [
        ("key_pm", _c_vp), ("key_tiles", _c_vp), ("key_maxnorm", _c_vp), ("val_pm", _c_vp),
        ("capacity_pos", _c_i64),
        ("K", _c_i32), ("CK", _c_i32), ("CV", _c_i32), ("val_dtype", _c_i32),
    ]


MAX_RANKS = 16


class Peers(ctypes.Structure):
    """struct EvavosPeers (include/evavos.h)."""
    _fields_ = [("n_ranks", _c_i32), ("rank", _c_i32), ("base", _c_vp * MAX_RANKS)]


class MemReadArgs(ctypes.Structure):
    """struct EvavosMemReadArgs (include/evavos.h)."""
    _fields_ = [
        ("bank", BankShadow),
        ("query", _c_vp), ("readout", _c_vp), ("topk_idx", _c_vp), ("topk_weight", _c_vp),
        ("topk_score", _c_vp), ("workspace", _c_vp),
        ("workspace_bytes", _c_i64), ("n_pos", _c_i64), ("n_query", _c_i64), ("query_ch_stride", _c_i64),
        ("readout_obj_stride", _c_i64), ("readout_ch_stride", _c_i64),
        ("top_k", _c_i32), ("path", _c_i32), ("n_sm", _c_i32), ("sample_stride", _c_i32),
        ("peers", ctypes.POINTER(Peers)), ("peer_gather_offset", _c_i64),
        ("queries_per_frame", _c_i64), ("readout_frame_stride", _c_i64),
    ]


# name -> (restype, argtypes); every symbol include/evavos.h declares.
SIGNATURES = {
    "evavos_abi_version": (_c_i32, []),
    "evavos_last_error": (ctypes.c_char_p, []),
    "evavos_sizeof_bank_shadow": (ctypes.c_size_t, []),
    "evavos_sizeof_memread_args": (ctypes.c_size_t, []),
    "evavos_key_tiles_bytes": (ctypes.c_size_t, [_c_i64]),
    "evavos_bank_write_keys": (_c_i32, [ctypes.POINTER(BankShadow), _c_vp, _c_i64, _c_i64, _c_i64, _c_vp, _c_i64, _c_vp]),
    "evavos_bank_write_values": (_c_i32, [ctypes.POINTER(BankShadow), _c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp,
                                          _c_i64, _c_i64, _c_vp]),
    "evavos_memread_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(MemReadArgs)]),
    "evavos_memread": (_c_i32, [ctypes.POINTER(MemReadArgs), _c_vp]),
    "evavos_memread_overflow_count": (_c_i32, [ctypes.POINTER(MemReadArgs), ctypes.POINTER(ctypes.c_uint32), _c_vp]),
    "evavos_readout": (_c_i32, [ctypes.POINTER(BankShadow), _c_vp, _c_vp, _c_i64, _c_i32, _c_vp, _c_i64, _c_i64, _c_vp]),
    "evavos_readout_qmajor": (_c_i32, [ctypes.POINTER(BankShadow), _c_vp, _c_vp, _c_i64, _c_i32, _c_vp, _c_vp]),
    "evavos_peer_enable": (_c_i32, [_c_i32]),
    "evavos_peer_buffer_alloc": (_c_i32, [_c_i64, ctypes.POINTER(_c_vp), ctypes.c_char_p]),
    "evavos_peer_buffer_open": (_c_i32, [ctypes.c_char_p, ctypes.POINTER(_c_vp)]),
    "evavos_peer_buffer_close": (_c_i32, [_c_vp]),
    "evavos_peer_buffer_free": (_c_i32, [_c_vp]),
    "evavos_peer_barrier": (_c_i32, [ctypes.POINTER(Peers), _c_i64, ctypes.c_uint32, _c_vp]),
    "evavos_peer_reduce_scatter": (_c_i32, [ctypes.POINTER(Peers), _c_i64, _c_i32, _c_i64, _c_i64, _c_vp, _c_i64, _c_vp]),
    "evavos_jf_workspace_bytes": (ctypes.c_size_t, [_c_i64, _c_i32, _c_i32]),
    "evavos_jf_metrics": (_c_i32, [_c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_i32, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp]),
    "evavos_affinity_dense": (_c_i32, [_c_vp, _c_vp, _c_i64, _c_i32, _c_i64, _c_vp, _c_vp]),
    "evavos_aggregate_wbg": (_c_i32, [_c_vp, _c_vp, _c_i32, _c_i64, _c_i32, _c_i32, _c_vp]),
    "evavos_bias_residual_nhwc": (_c_i32, [_c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_i32, _c_vp]),
    "evavos_upsample2x_add_nhwc": (_c_i32, [_c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_i32, _c_i32, _c_vp]),
    "evavos_argmax_unpad": (_c_i32, [_c_vp, _c_i32, _c_i64, _c_i32, _c_i32, _c_vp, _c_vp, _c_i32, _c_i32, _c_i32, _c_i32, _c_vp]),
    "evavos_attention_workspace_bytes": (ctypes.c_size_t, [_c_i32, _c_i64, _c_i64, _c_i32]),
    "evavos_attention_readout": (_c_i32, [_c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_i64, _c_i32, _c_i32, _c_i64, _c_i64,
                                          _c_vp, _c_i64, _c_vp, _c_i64, _c_i32, _c_vp]),
    "evavos_topk_merge": (_c_i32, [_c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_i32, _c_i32, _c_i64, _c_vp, _c_vp, _c_vp,
                                   _c_vp, _c_vp]),
    "evavos_topk_merge_gathered": (_c_i32, [_c_vp, _c_i64, _c_i32, _c_i32, _c_i32, _c_i32, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp,
                                            _c_vp]),
    "evavos_memread_host": (_c_i32, [_c_vp, _c_vp, _c_vp, _c_i32, _c_i32, _c_i32, _c_i64, _c_i64, _c_i32, _c_i32,
                                     _c_vp, _c_vp, _c_vp, ctypes.POINTER(_c_i64), ctypes.POINTER(_c_i64)]),
    "evavos_release_host_scratch": (_c_i32, []),
    "evavos_stage_timing": (_c_i32, [_c_i32]),
    "evavos_stage_timing_read": (_c_i32, [ctypes.POINTER(ctypes.c_float)]),
}


class EvavosError(RuntimeError):
    """A C-ABI call returned a negative status; carries the code and evavos_last_error()."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libevavos_sm100 error {code}: {message}")
        self.code = code


_lock = threading.Lock()
_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library once; raise loudly if it is missing or mismatched."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the sm_100a CUDA library has not been built. "
                "Run `python -m evavos_b200.build`. There is no CPU or PyTorch fallback for this path.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.evavos_abi_version() != ABI_VERSION:
            raise ImportError(f"{LIB_NAME}: ABI {lib.evavos_abi_version()} != expected {ABI_VERSION}; rebuild")
        if lib.evavos_sizeof_bank_shadow() != ctypes.sizeof(BankShadow) or \
                lib.evavos_sizeof_memread_args() != ctypes.sizeof(MemReadArgs):
            raise ImportError(f"{LIB_NAME}: struct layout mismatch between include/evavos.h and _lib.py")
        _lib = lib
    return _lib


def check(code: int) -> None:
    if code != 0:
        msg = load().evavos_last_error()
        raise EvavosError(code, msg.decode("utf-8", "replace") if msg else "")


def current_stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream
