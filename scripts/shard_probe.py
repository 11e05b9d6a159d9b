"""Single-GPU emulation of one rank's work in the 8-way sharded cfg4 read (no collectives): where does the time go?"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev
from evavos_b200 import _lib
from evavos_b200.sharded import CudaShardOps
dev = torch.device("cuda:0")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ck, cv, h, w, k, F = 64, 512, 30, 54, 1, 5
t_local = 200 // world
g = torch.Generator().manual_seed(1)
bank = ev.MemoryBank(k, ck, cv, h, w, t_local, dev, keep_reference_layout=False)
for f in range(t_local):
    bank.append(torch.randn(1, ck, h, w, generator=g).to(dev), torch.randn(k, cv, 1, h, w, generator=g).to(dev))
qk = torch.randn(1, ck, F, h, w, generator=g).to(dev)
ops = CudaShardOps()
def ev_time(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
idx, sc = ops.local_topk(bank, qk, 50)
packed = torch.stack([idx, sc.view(torch.int32)], -1).contiguous()
gathered = packed.unsqueeze(0).repeat(world, 1, 1, 1).contiguous()
gi, wt, loc = ops.merge_gathered(gathered, 50, 0, world, h * w)
print(f"world={world} local frames={t_local} queries={F*h*w}")
print("local_topk      us", ev_time(lambda: ops.local_topk(bank, qk, 50)))
print("stack           us", ev_time(lambda: torch.stack([idx, sc.view(torch.int32)], -1).contiguous()))
print("merge_gathered  us", ev_time(lambda: ops.merge_gathered(gathered, 50, 0, world, h * w)))
print("readout         us", ev_time(lambda: ops.readout(bank, loc, wt)))
lib = _lib.load(); lib.evavos_stage_timing(1)
acc = np.zeros(4)
for i in range(13):
    ops.local_topk(bank, qk, 50)
    ms = (ctypes.c_float * 4)(); lib.evavos_stage_timing_read(ms)
    if i >= 3: acc += np.array(list(ms))
print("local_topk stages (us): filter %.1f | fallback %.1f | finalize %.1f" % tuple(acc[:3] / 10 * 1e3))
