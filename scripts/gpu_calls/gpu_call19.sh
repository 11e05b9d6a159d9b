mkdir -p gpurun_out/r2
timeout 300 ncu --set full --import-source on --clock-control none -k regex:score_select -s 3 -c 1 -f -o gpurun_out/r2/full_cfg2x5 python scripts/profile_batched.py cfg2 > gpurun_out/r2/ncu_cfg2x5.log 2>&1; echo "ncu rc=$?"
for R in 1 2; do echo "R=$R"; EVAVOS_SAMPLE_STRIDE=$R python - <<'PY'
import os, sys, ctypes, numpy as np, torch
sys.path.insert(0, os.getcwd())
import evavos_b200 as ev
from evavos_b200 import _lib
from bench import WORKLOADS, TOP_K
dev=torch.device('cuda:0'); lib=_lib.load()
ck,cv,t,h,w,k,seed,_=WORKLOADS['cfg2']
g=torch.Generator(device=dev).manual_seed(seed)
bank=ev.MemoryBank(1,ck,cv,h,w,t,dev,keep_reference_layout=False)
bank.write_frames(0, torch.randn(1,ck,t,h,w,generator=g,device=dev), torch.randn(1,cv,t,h,w,generator=g,device=dev))
qk=torch.randn(1,ck,5,h,w,generator=g,device=dev)
lib.evavos_stage_timing(1)
for i in range(23): ev.memory_read(bank,qk,TOP_K)
ms=(ctypes.c_float*4)(); lib.evavos_stage_timing_read(ms); print('cfg2 x5 frames: filter %.1f finalize %.1f readout %.1f us' % (ms[0]*1e3, ms[2]*1e3, ms[3]*1e3))
PY
done
