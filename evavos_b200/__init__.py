"""evavos_b200 - the STCN/MiVOS space-time memory read of EVA-VOS on B200 (sm_100a).

Host code is Python/PyTorch (device memory, streams, torch.distributed); the arithmetic is in
hand-written CUDA behind the C ABI of include/evavos.h (libevavos_sm100.so, loaded with ctypes).
Importing the package does not load the library; the first kernel call does, and fails loudly
if it has not been built (no CPU fallback).
"""
from . import _lib
from ._lib import EvavosError
from .aggregate import aggregate_wbg, argmax_unpad, get_segmentations
from .attention import attention_readout
from .memory_bank import MemoryBank
from .memory_reader import EvalMemoryReader, TopKAffinity, memory_read
from .metrics import eval_processor_metric, frame_metrics, get_j_and_f
from .networks import FusionNet, PropagationNetwork
from .inference_core import InferenceCore
from .tensor_util import pad_divide_by

__all__ = ["EvavosError", "aggregate_wbg", "argmax_unpad", "get_segmentations", "attention_readout", "MemoryBank", "EvalMemoryReader", "TopKAffinity", "memory_read",
           "PropagationNetwork", "FusionNet", "InferenceCore", "pad_divide_by", "_lib",
           "eval_processor_metric", "frame_metrics", "get_j_and_f"]
