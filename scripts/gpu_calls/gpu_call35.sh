#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_sharded.py tests/test_gpu_inference_core.py -x -q -m gpu 2>&1 | tail -3
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import evavos_b200 as ev
from evavos_b200 import _lib
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(5)
h, w = 30, 54
for t in (1, 3, 5, 8):
    base = torch.randn(1, 64, 1, 1, 1, generator=g, device=dev)
    bank = ev.MemoryBank(1, 64, 512, h, w, t, dev, keep_reference_layout=False)
    bank.write_frames(0, base + 1e-3 * torch.randn(1, 64, t, h, w, generator=g, device=dev), torch.randn(1, 512, t, h, w, generator=g, device=dev))
    qk = 0.9 * base + 1e-3 * torch.randn(1, 64, 5, h, w, generator=g, device=dev)
    for _ in range(3):
        ev.memory_read(bank, qk, 50, path=_lib.PATH_TENSOR_DENSE, want_readout=False, want_topk=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ev.memory_read(bank, qk, 50, path=_lib.PATH_TENSOR_DENSE, want_readout=False, want_topk=True)
    e1.record(); torch.cuda.synchronize()
    from evavos_b200.memory_reader import last_overflow_count
    print(f"{t * h * w} positions x 8100 queries, {last_overflow_count()} overflowed: filter + finalize + exact pass {1e3 * e0.elapsed_time(e1) / 10:.0f} us")
PY
python scripts/profile_cfg3.py cl graphs 2>&1 | head -1
python scripts/profile_cfg3.py amp graphs 2>&1 | head -1
exit 0
