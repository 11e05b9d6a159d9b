"""Warm stage times of the memory read (filter / exact fallback / finalize / readout) per workload.

    [EVAVOS_LIB=path/to/variant.so] python scripts/filter_time.py [cfg1 cfg2 cfg4 ...]
Used to time compile-time variants of the score filter (-DEVAVOS_EXP=...) against the production build.
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from evavos_b200 import _lib  # noqa: E402
from bench import WORKLOADS, synth  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
for name in (sys.argv[1:] or ["cfg1", "cfg2", "cfg4"]):
    ck, cv, t, h, w, k, seed, _ = WORKLOADS[name]
    k = int(os.environ.get("FILTER_K", k))      # the filter does not depend on the values: FILTER_K=1 saves host time
    mk, qk, mv = synth(seed, ck, cv, t, h, w, k)
    bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, keep_reference_layout=False)
    bank.write_frames(0, mk.to(dev), mv.to(dev))
    qk = qk.to(dev)
    lib.evavos_stage_timing(1)
    acc, reps = np.zeros(4), 20
    for i in range(reps + 3):
        ev.memory_read(bank, qk, 50, path=int(os.environ.get("FILTER_PATH", 0)))
        ms = (ctypes.c_float * 4)()
        lib.evavos_stage_timing_read(ms)
        if i >= 3:
            acc += np.array(list(ms))
    lib.evavos_stage_timing(0)
    us = acc / reps * 1e3
    print(f"[{name}] filter {us[0]:.1f} | exact fallback {us[1]:.1f} | finalize {us[2]:.1f} | readout {us[3]:.1f} us", flush=True)
    del bank
