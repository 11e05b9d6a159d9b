import ctypes, os, sys
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
os.environ["EVAVOS_LIB"] = os.path.join(root, "evavos_b200", "libevavos_sm100_trace.so")
import evavos_b200 as ev
from evavos_b200 import _lib
from bench import WORKLOADS, synth
ck, cv, t, h, w, k, seed, _ = WORKLOADS["cfg2"]
dev = torch.device("cuda:0")
mk, qk, mv = synth(seed, ck, cv, t, h, w, k)
bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, keep_reference_layout=False)
bank.write_frames(0, mk.to(dev), mv.to(dev)); qk = qk.to(dev)
lib = _lib.load(); lib.evavos_stage_timing(1)
for flags in (0,):
    lib.evavos_debug_fin_skip(flags)
    acc = np.zeros(4)
    for i in range(13):
        ev.memory_read(bank, qk, 50)
        ms = (ctypes.c_float * 4)(); lib.evavos_stage_timing_read(ms)
        if i >= 3: acc += np.array(list(ms))
    print(f"fin_skip={flags}: finalize {acc[2] / 10 * 1e3:.1f} us (filter {acc[0] / 10 * 1e3:.1f}, fallback {acc[1] / 10 * 1e3:.1f})", flush=True)
