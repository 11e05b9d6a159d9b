"""Does one step (fused read + aggregate) capture into a CUDA graph, and what does replay cost per step?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev
from bench import WORKLOADS, synth, TOP_K
ck, cv, t, h, w, k, seed, _ = WORKLOADS["cfg2"]
dev = torch.device("cuda:0")
banks, qs = [], []
for b in range(4):
    mk, qk, mv = synth(seed + b, ck, cv, t, h, w, k)
    bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, keep_reference_layout=False)
    bank.write_frames(0, mk.to(dev), mv.to(dev))
    banks.append(bank); qs.append(qk.to(dev))
prob = torch.rand(k, 1, h * 16, w * 16, device=dev)
def step(b):
    out, _ = ev.memory_read(banks[b], qs[b], TOP_K)
    return out, ev.aggregate_wbg(prob, keep_bg=True)
for b in range(4):
    step(b)
torch.cuda.synchronize()
graphs, keep = [], []
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for b in range(4):
        step(b)
torch.cuda.synchronize()
for b in range(4):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep.append(step(b))
    graphs.append(g)
ref = [step(b) for b in range(4)]
for b in range(4):
    graphs[b].replay()
torch.cuda.synchronize()
for b in range(4):
    assert torch.equal(keep[b][0], ref[b][0]) and torch.equal(keep[b][1], ref[b][1])
def timeit(fn, n=200):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("eager  us/step", timeit(lambda i: step(i % 4)))
print("graph  us/step", timeit(lambda i: graphs[i % 4].replay()))
