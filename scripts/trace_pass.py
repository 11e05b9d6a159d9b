"""Print the in-kernel timeline of CTA 0 of the last score pass (needs the -DEVAVOS_TRACE build)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("EVAVOS_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "evavos_b200", "libevavos_sm100_trace.so"))
import evavos_b200 as ev
from evavos_b200 import _lib
from bench import WORKLOADS, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
ck, cv, t, h, w, k, seed, _ = WORKLOADS[name]
dev = torch.device("cuda:0")
mk, qk, mv = synth(seed, ck, cv, t, h, w, k)
bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, keep_reference_layout=False)
bank.write_frames(0, mk.to(dev), mv.to(dev)); qk = qk.to(dev)
nf = None
if len(sys.argv) > 2 and sys.argv[2] == "nomma":
    # odd n_pos triggers the trace build's "skip MMAs" experiment: read one position less (partial last tile)
    import evavos_b200.memory_reader as mr
    _orig = mr.memory_read
for _ in range(2):
    if len(sys.argv) > 2 and sys.argv[2] == "nomma":
        a = _lib.MemReadArgs()
        a.bank = bank.shadow(); q2 = qk.reshape(ck, -1)
        a.query = q2.data_ptr(); a.query_ch_stride = q2.stride(0)
        a.n_pos, a.n_query, a.top_k, a.path = bank.n_pos - 1, q2.shape[1], 50, 1
        idx = torch.empty((q2.shape[1], 50), dtype=torch.int32, device=dev); wt = torch.empty((q2.shape[1], 50), device=dev)
        a.topk_idx, a.topk_weight = idx.data_ptr(), wt.data_ptr()
        a.n_sm = 148
        lib0 = _lib.load()
        need = lib0.evavos_memread_workspace_bytes(ctypes.byref(a))
        ws = torch.empty((need + 4096,), dtype=torch.uint8, device=dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        print("memread rc", lib0.evavos_memread(ctypes.byref(a), torch.cuda.current_stream().cuda_stream))
    else:
        ev.memory_read(bank, qk, 50, want_readout=False, want_topk=True)
torch.cuda.synchronize()
lib = _lib.load()
i0 = int(os.environ.get("TRACE_I0", "0"))
if i0:
    lib.evavos_debug_trace_base.argtypes = [ctypes.c_int]
    lib.evavos_debug_trace_base(i0)
    ev.memory_read(bank, qk, 50, want_readout=False, want_topk=True)
    torch.cuda.synchronize()
print("trace window starts at iteration", i0)
buf = (ctypes.c_longlong * (6 * 64))()
lib.evavos_debug_trace.argtypes = [ctypes.c_void_p]
print("rc", lib.evavos_debug_trace(buf))
tr = np.array(list(buf), dtype=np.int64).reshape(6, 64)
t0 = tr[1, 0] if i0 == 0 else tr[0, 57]
names = ["full_ready", "mma_ready", "mma_issued", "epi_accfull", "epi_ldtm_done", "epi_math_done"]
print("tile " + " ".join(f"{n:>13s}" for n in names))
for i in range(0, 40):
    print(f"{i:4d} " + " ".join(f"{int(tr[r, i] - t0):13d}" for r in range(6)))
print("kernel marks (entry, roles start, end sweep2, exit):", [int(x - t0) for x in tr[0, 56:60]])
print("phase marks (end phase A, after barrier1, after thresholds, after barrier2):", [int(x - t0) for x in tr[0, 60:64]])
d = np.diff(tr[:, 8:40], axis=1)
print("mean cycles/tile (tiles 8..40):", {n: float(d[r].mean()) for r, n in enumerate(names)})
print("issued(i-1) -> full_ready(i)", float((tr[0, 9:40] - tr[2, 8:39]).mean()), " full_ready -> mma_ready (wait for acc_empty)", float((tr[1, 8:40] - tr[0, 8:40]).mean()))
print("mma_ready -> issued", float((tr[2, 8:40] - tr[1, 8:40]).mean()), " accfull -> ldtm", float((tr[4, 8:40] - tr[3, 8:40]).mean()),
      " ldtm -> math", float((tr[5, 8:40] - tr[4, 8:40]).mean()), " issued(i) -> accfull(i)", float((tr[3, 8:40] - tr[2, 8:40]).mean()))
