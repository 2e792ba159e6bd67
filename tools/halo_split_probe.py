#!/usr/bin/env python
"""Time the stride-1 3x3 tensor-core layers alone, with and without the split-reduction workspace (not a test).
    python tools/halo_split_probe.py [math]       (math: 1 = 3xTF32 parity, 3 = tf32, 4 = bf16)
Each layer: CUDA-graph of 20 back-to-back launches (the ctypes launch path costs more host time than the kernel
takes), L2 warm; and single launches with the L2 flushed in between."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from robocupvision_b200 import ops
from umma_probe import timeit

math = int(sys.argv[1]) if len(sys.argv) > 1 else 1
flush = torch.empty(64 * 1024 * 1024, device="cuda")


def cold(fn, n=12):
    ts = []
    for i in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for cin, cout, dil, n, h, w in [(128, 128, 1, 64, 15, 20), (64, 64, 1, 64, 15, 20), (128, 64, 1, 64, 15, 20),
                                (64, 128, 1, 64, 15, 20), (32, 32, 1, 64, 30, 40), (128, 128, 1, 256, 15, 20),
                                (128, 128, 1, 1, 15, 20), (128, 128, 2, 1, 15, 20), (128, 128, 1, 8, 15, 20),
                                (64, 64, 1, 1, 15, 20), (32, 32, 1, 1, 30, 40),
                                (128, 128, 2, 8, 30, 40), (64, 128, 2, 8, 30, 40), (128, 128, 2, 1, 30, 40),
                                (128, 128, 1, 32, 15, 20), (128, 128, 1, 40, 15, 20)]:
    g = ops.ConvGeom(cin, cout, 3, 1, dil, dil, False)
    x = torch.randn(n, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, 3, 3, device="cuda") / (cin * 9) ** 0.5
    y = torch.empty(n, cout, h, w, device="cuda")
    wp = ops.conv_pack(g, wt, ops.PACK_FWD, math=math, nhw=(n, h, w))
    need = ops.conv_workspace_bytes(g, n, h, w, ops.PACK_FWD, math)
    ws = ops.new_workspace(need, "cuda") if need else None
    row = []
    for wsp in (None, ws):
        fn = lambda: ops.conv_fwd(g, x, wt, None, epilogue=ops.EPI_RELU, out=y, math=math, wpacked=wp, workspace=wsp)
        row.append((timeit(fn), cold(fn)))
    fl = 2.0 * cin * cout * 9 * h * w * n
    print(f"{cin:3d}->{cout:3d} d{dil} {n:3d}x{h}x{w} math {math}: whole {row[0][0]:6.1f} us warm / {row[0][1]:6.1f} cold | "
          f"split {row[1][0]:6.1f} / {row[1][1]:6.1f} (workspace {need / 1e6:.2f} MB)  "
          f"{fl / (min(row[0][0], row[1][0]) * 1e-6) / 1e12:6.1f} TFLOP/s", flush=True)
