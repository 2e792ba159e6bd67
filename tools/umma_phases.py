#!/usr/bin/env python
"""Phase timing (clock64 samples of CTA 0) of the tensor-core conv kernel on one layer (debug tool)."""
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from robocupvision_b200 import _lib, ops

mode = "wgrad" if "wgrad" in sys.argv else "fwd"
nums = [a for a in sys.argv[1:] if a.isdigit()]
cin, cout, h, w, B = (int(a) for a in nums[:5]) if len(nums) >= 5 else (128, 128, 15, 20, 64)
lib = _lib.load()
lib.rcv_debug_set_prof.argtypes = [C.c_void_p]
lib.rcv_debug_set_prof.restype = None
geo = ops.ConvGeom(cin, cout, 3, 1, 1, 1, False)
x = torch.randn(B, cin, h, w, device="cuda"); wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
wp = ops.conv_pack(geo, wt, ops.PACK_FWD)
dy = torch.randn(B, cout, h, w, device="cuda"); dw = torch.zeros_like(wt)
run = (lambda: ops.conv_wgrad(geo, x, dy, dw=dw, math=ops.MATH_TF32X3)) if mode == "wgrad" else \
      (lambda: ops.conv_fwd(geo, x, wt, None, math=ops.MATH_TF32X3, wpacked=wp))
for _ in range(3):
    run()
prof = torch.zeros(8192, dtype=torch.int64, device="cuda")
lib.rcv_debug_set_prof(C.c_void_p(prof.data_ptr()))
run()
torch.cuda.synchronize()
lib.rcv_debug_set_prof(None)
pr = prof.cpu().numpy()
nkb = (cin * 9 + 31) // 32 if mode == 'fwd' else 48
nkb = max([kb for kb in range(128) if pr[kb * 8]] + [0]) + 1
t0 = min(int(pr[kb * 8]) for kb in range(min(nkb, 128)) if pr[kb * 8])
print("producer (row 0 of each group): kb: start loads_issued(+) wait_empty(+) stores(+) fence(+) arrive(+)")
for kb in range(min(nkb, 128)):
    e = [int(v) for v in pr[kb * 8: kb * 8 + 6]]
    print(f"  kb {kb:3d}: {e[0]-t0:7d} " + " ".join(f"+{e[i+1]-e[i]:5d}" for i in range(5)))
print("issuer: kb: start  mma_issue(+) next_full_wait(+) commits(+)")
for kb in range(min(nkb, 128)):
    e = [int(v) for v in pr[2048 + kb * 4: 2048 + kb * 4 + 4]]
    print(f"  kb {kb:3d}: {e[0]-t0:7d} +{e[1]-e[0]:5d} +{e[2]-e[1]:5d} +{e[3]-e[2]:5d}")
for g in range(4):
    e = [int(v) for v in pr[4000 + g * 4: 4000 + g * 4 + 3]]
    if e[0]:
        print(f"group {g}: loop end {e[0]-t0}, done wait +{e[1]-e[0]}, epilogue +{e[2]-e[1]}")
