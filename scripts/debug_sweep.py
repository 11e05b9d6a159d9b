"""Timing experiment: pass kernels under EVAVOS_DEBUG flags (results invalid, timing only)."""
import os, sys, subprocess
code = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
import evavos_b200 as ev
from bench import WORKLOADS, synth
ck, cv, t, h, w, k, seed, _ = WORKLOADS[sys.argv[1]]
dev = torch.device("cuda:0")
mk, qk, mv = synth(seed, ck, cv, t, h, w, k)
bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, keep_reference_layout=False)
bank.write_frames(0, mk.to(dev), mv.to(dev)); qk = qk.to(dev)
def run(): ev.memory_read(bank, qk, 50, want_readout=False, want_topk=True)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"{sys.argv[1]} EVAVOS_DEBUG={os.environ.get('EVAVOS_DEBUG','0')}: select {e0.elapsed_time(e1)*100:.0f} us", flush=True)
'''
for cfg in ("cfg4",):
    for flags in (0, 1, 2, 9, 5, 13, 29, 31):
        env = dict(os.environ, EVAVOS_DEBUG=str(flags))
        subprocess.run([sys.executable, "-c", code, cfg], env=env)
