"""Quick on-GPU diagnostics: correctness of both selection paths vs the oracle and coarse timings.

    python tests/gpu_probe.py [--skip-tensor]
Checker tooling (it uses the oracle, so it lives under tests/); not a benchmark (bench.py is); prints enough detail to debug a failing path from one gpurun call.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev  # noqa: E402
from evavos_b200 import _lib  # noqa: E402
from oracle import memread_np as onp  # noqa: E402
from tests.helpers import TIE_TOL, synth  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters * 1e3  # us


def check(tag, mk, qk, mv, top_k, path, s64=None):
    dev = torch.device("cuda:0")
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    out, aff = ev.memory_read(bank, qk.to(dev), top_k, want_topk=True, path=path)
    torch.cuda.synchronize()
    idx = aff.idx.cpu().numpy()
    w = aff.weight.cpu().numpy()
    ck = mk.shape[1]
    if s64 is None:
        s64 = onp.affinity_scores(mk[0].reshape(ck, -1).numpy(), qk[0].reshape(ck, -1).numpy())
    exact, tie, bad, bad_q = onp.compare_topk(idx, s64, top_k, TIE_TOL)
    tk = onp.topk_softmax(s64, top_k)
    ro = onp.readout(tk.idx, tk.weight, mv.reshape(mv.shape[0], mv.shape[1], -1).numpy())
    err = onp.rel_l2(out.cpu().numpy().reshape(ro.shape), ro)
    print(f"[{tag}] path={path} exact={exact} tie={tie} bad={bad} readout rel-L2={err:.3e} "
          f"wsum_err={np.abs(w.sum(1) - 1).max():.2e}", flush=True)
    if bad:
        q = bad_q[0]
        print("   first bad query", q, "got", np.sort(idx[q])[:12], "want", np.sort(tk.idx[q])[:12], flush=True)
    return s64, bad == 0 and err < 1e-3


def main():
    skip_tensor = "--skip-tensor" in sys.argv
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs", flush=True)
    ok = True
    cases = [("tiny", synth(11, 64, 32, 3, 6, 9, 2), 50), ("cfg1", synth(1235, 64, 512, 5, 30, 54, 1), 50)]
    for tag, (mk, qk, mv), k in cases:
        s64, good = check(tag, mk, qk, mv, k, _lib.PATH_SIMT)
        ok &= good
        if not skip_tensor:
            _, good = check(tag, mk, qk, mv, k, _lib.PATH_TENSOR, s64)
            ok &= good
    # timings
    dev = torch.device("cuda:0")
    for tag, (ck, cv, t, h, w, k) in (("cfg1", (64, 512, 5, 30, 54, 1)), ("cfg2", (64, 512, 20, 30, 54, 3)),
                                      ("cfg4", (64, 512, 200, 30, 54, 1))):
        mk, qk, mv = synth(1234, ck, cv, t, h, w, k)
        mk, qk, mv = mk.to(dev), qk.to(dev), mv.to(dev)
        t0 = time.time()
        bank = ev.MemoryBank.from_tensors(mk, mv)
        torch.cuda.synchronize()
        print(f"[{tag}] shadow build {1e3 * (time.time() - t0):.1f} ms (first call incl. alloc)", flush=True)
        us_build = timed(lambda: bank.write_frames(0, mk, mv), iters=5, warm=1)
        line = f"[{tag}] bank import {us_build:.0f} us"
        for name, path in (("simt", _lib.PATH_SIMT), ("tensor", _lib.PATH_TENSOR)):
            if path == _lib.PATH_TENSOR and skip_tensor:
                continue
            if path == _lib.PATH_SIMT and tag == "cfg4":
                continue
            us = timed(lambda: ev.memory_read(bank, qk, 50, path=path), iters=10, warm=2)
            us_sel = timed(lambda: ev.memory_read(bank, qk, 50, want_readout=False, want_topk=True, path=path),
                           iters=10, warm=2)
            line += f" | {name}: read {us:.0f} us (select only {us_sel:.0f} us)"
        print(line, flush=True)
        if not skip_tensor:
            import ctypes
            lib = _lib.load()
            lib.evavos_stage_timing(1)
            acc = np.zeros(4)
            reps = 20
            for i in range(reps + 3):
                ev.memory_read(bank, qk, 50)
                ms = (ctypes.c_float * 4)()
                lib.evavos_stage_timing_read(ms)
                if i >= 3:
                    acc += np.array(list(ms))
            lib.evavos_stage_timing(0)
            print(f"[{tag}] warm stage times (us): filter {acc[0] / reps * 1e3:.1f} | exact fallback {acc[1] / reps * 1e3:.1f} | "
                  f"finalize {acc[2] / reps * 1e3:.1f} | readout {acc[3] / reps * 1e3:.1f}", flush=True)
        del bank
    p = torch.rand(3, 1, 480, 864, device=dev)
    us = timed(lambda: ev.aggregate_wbg(p, keep_bg=True), iters=50)
    print(f"[aggregate K=3 480x864] {us:.1f} us -> {(7 * 480 * 864 * 4) / us / 1e3:.0f} GB/s", flush=True)
    print("PROBE_OK" if ok else "PROBE_FAIL", flush=True)


if __name__ == "__main__":
    main()
