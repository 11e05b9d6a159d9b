"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evavos_b200 as ev
from evavos_b200 import _lib
from evavos_b200.sharded import CudaShardOps, local_to_global
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
K, CK, CV, T, H, W = 2, 64, 512, 3, 9, 13          # N = 351 (partial tiles), HW = 117 (partial query tile)
mk = torch.randn(1, CK, T, H, W, generator=g).to(dev)
mv = torch.randn(K, CV, T, H, W, generator=g).to(dev)
qk = torch.randn(1, CK, H, W, generator=g).to(dev)
bank = ev.MemoryBank(K, CK, CV, H, W, T + 1, dev)
for f in range(T):
    bank.append(mk[:, :, f], mv[:, :, f:f + 1])
for path in (_lib.PATH_TENSOR, _lib.PATH_SIMT):
    out, aff = ev.memory_read(bank, qk, 50, want_topk=True, path=path)
    dense = aff.to_dense()
bank16 = ev.MemoryBank.from_tensors(mk, mv, value_dtype=torch.bfloat16)
out16, _ = ev.memory_read(bank16, qk, 50)
small = ev.MemoryBank.from_tensors(mk[:, :32].contiguous(), mv[:, :24].contiguous())   # CK=32, generic readout
outs, _ = ev.memory_read(small, qk[:, :32].contiguous(), 20)
agg = ev.aggregate_wbg(torch.rand(3, 1, 33, 47, device=dev), keep_bg=True)
agg11 = ev.aggregate_wbg(torch.rand(11, 1, 16, 16, device=dev), keep_bg=False, hard=True)
ops = CudaShardOps()
idx, sc = ops.local_topk(bank, qk, 50)
packed = torch.stack([idx, sc.view(torch.int32)], -1).unsqueeze(0).contiguous()
gi, w, loc = ops.merge_gathered(packed, 50, 0, 1, H * W)
part = ops.readout(bank, loc, w)
att = ev.attention_readout(mk[:, :, 1:2], qk, torch.rand(6, H * W, generator=g).to(dev))      # fusion-path attention read
masks, unp = ev.argmax_unpad(torch.rand(3, 4, 1, 32, 48, device=dev), (3, 5, 2, 6), 24, 40)
torch.cuda.synchronize()
print("sanitize run ok", float(out.abs().mean()), float(part.abs().mean()))
