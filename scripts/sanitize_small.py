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
# round 2: multi-frame query batch into a frame-major destination, the overflow pass (near-constant keys, bank length not
# a multiple of 128), query-major readout, J&F scoring, the peer barrier / reduce-scatter on a single rank
qk3 = torch.randn(1, CK, 3, H, W, generator=g).to(dev)
m4 = torch.empty((3, K, 2 * CV, H, W), device=dev)
ev.memory_read(bank, qk3, 50, out=m4)
base = torch.randn(1, CK, 1, 1, 1, generator=g)
Td, Hd, Wd = 2, 24, 30                                                                  # 1 440 positions
dense_bank = ev.MemoryBank.from_tensors((base + 1e-3 * torch.randn(1, CK, Td, Hd, Wd, generator=g)).to(dev),
                                        torch.randn(1, 64, Td, Hd, Wd, generator=g).to(dev))
qd = (0.9 * base[:, :, 0] + 1e-3 * torch.randn(1, CK, 5, 7, generator=g)).to(dev)
for p_ in (_lib.PATH_TENSOR_DENSE, _lib.PATH_TENSOR):
    out_d, aff_d = ev.memory_read(dense_bank, qd, 50, want_topk=True, path=p_)
from evavos_b200.memory_reader import last_overflow_count
n_over = last_overflow_count()
pred = torch.rand(3, 37, 53, generator=g).to(dev) > 0.5
gt = torch.rand(3, 37, 53, generator=g).to(dev) > 0.5
jf = ev.frame_metrics(pred, gt)
# round 2, later: the tensor-core form of the attention read (ragged tiles, two split counts, 4 / 8 / 16-row variants)
# and the decoder's elementwise tails (fp32 and bf16)
os.environ["EVAVOS_ATTENTION_PATH"] = "tensor"
for rows in (3, 6, 19):
    att_tc = ev.attention_readout(mk[:, :, 1:2], qk, torch.rand(rows, H * W, generator=g).to(dev))
big_k, big_q = torch.randn(1, CK, 1, 20, 27, generator=g).to(dev), torch.randn(1, CK, 20, 27, generator=g).to(dev)
att_big = ev.attention_readout(big_k, big_q, torch.rand(4, 540, generator=g).to(dev))      # 8.4 tiles over several splits
os.environ.pop("EVAVOS_ATTENTION_PATH")
from evavos_b200.decoder_ops import bias_residual_, upsample2x_add_
for dt in (torch.float32, torch.bfloat16):
    y = torch.randn(2, 16, 6, 10, generator=g).to(dev, dt).contiguous(memory_format=torch.channels_last)
    x = torch.randn(2, 16, 3, 5, generator=g).to(dev, dt).contiguous(memory_format=torch.channels_last)
    bias_residual_(y, torch.randn(16, generator=g).to(dev), y.clone(memory_format=torch.preserve_format), relu=True)
    upsample2x_add_(y, torch.randn(16, generator=g).to(dev), x)
torch.cuda.synchronize()
print("sanitize run ok", float(out.abs().mean()), float(part.abs().mean()), "overflowed queries:", n_over)
