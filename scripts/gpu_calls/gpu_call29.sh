#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_sharded.py tests/test_gpu_memread.py -x -q -m gpu > gpurun_out/r2/c29_tests.txt 2>&1
tail -4 gpurun_out/r2/c29_tests.txt
python - <<'PY' > gpurun_out/r2/c29_merge.txt 2>&1
import torch, sys
sys.path.insert(0, '.')
import evavos_b200 as ev
from evavos_b200.sharded import CudaShardOps
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(1)
ops = CudaShardOps()
for world in (2, 8):
    nq, k = 8100, 50
    gathered = torch.stack([torch.randint(0, 25 * 1620, (world, nq, k), generator=g, device=dev).int(),
                            torch.randn(world, nq, k, generator=g, device=dev).sort(-1, descending=True).values.view(torch.int32)], -1).contiguous()
    for _ in range(3):
        ops.merge_gathered(gathered, k, 0, world, 1620)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.merge_gathered(gathered, k, 0, world, 1620)
    e1.record(); torch.cuda.synchronize()
    print(f"merge of {world} lists x {nq} queries: {1e3 * e0.elapsed_time(e1) / 20:.1f} us")
PY
cat gpurun_out/r2/c29_merge.txt
python scripts/profile_cfg3.py cl graphs 2>&1 | head -1
exit 0
