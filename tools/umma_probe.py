#!/usr/bin/env python
"""GPU probe: tcgen05 3xTF32 conv engine vs the ATen CPU op, error + time per geometry (not a test)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import torch.nn.functional as F
from robocupvision_b200 import ops

GEOMS = {"k3s1d1": (3, 1, 1, 1, False), "k3s1d2": (3, 1, 2, 2, False), "k3s2": (3, 2, 1, 1, False),
         "k1": (1, 1, 0, 1, False), "convT": (3, 2, 1, 1, True)}


def ref_conv(geom, x, w, b):
    k, s, p, d, tr = GEOMS[geom]
    if tr:
        return F.conv_transpose2d(x, w, b, stride=2, padding=1, output_padding=1)
    return F.conv2d(x, w, b, s, p, d)


def timeit(fn, n=20):
    """us per call, device time: n calls captured in a CUDA graph (the ctypes launch path costs
    ~35 us of host time per call, more than the short kernels take)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    quick = "--quick" in sys.argv
    if any(a.startswith("--layer=") for a in sys.argv):
        return timing()
    cases = [("k3s1d1", 40, 200, 2, 9, 11), ("k3s1d1", 8, 16, 2, 12, 20), ("k3s1d1", 128, 128, 2, 15, 20), ("k3s1d2", 64, 128, 2, 15, 20),
             ("k3s2", 16, 32, 2, 12, 20), ("convT", 64, 32, 2, 15, 20), ("k1", 16, 5, 2, 12, 20),
             ("k3s1d1", 5, 7, 2, 9, 7), ("k3s1d1", 24, 40, 3, 12, 20), ("k3s1d1", 3, 8, 2, 12, 20)]
    for geom, cin, cout, n, h, w in cases:
        k, s, p, d, tr = GEOMS[geom]
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n, cin, h, w, generator=g)
        wt = torch.randn((cin, cout, 3, 3) if tr else (cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5
        b = torch.randn(cout, generator=g)
        ref = ref_conv(geom, x, wt, b)
        geo = ops.ConvGeom(cin, cout, k, s, p, d, tr)
        for math, nm in ((ops.MATH_FP32, "fp32"), (ops.MATH_TF32X3, "tf32x3")):
            try:
                got = ops.conv_fwd(geo, x.cuda(), wt.cuda(), b.cuda(), math=math)
                torch.cuda.synchronize()
                err = float((got.cpu() - ref).abs().max()) / max(1.0, float(ref.abs().max()))
                print(f"fwd   {geom:7s} {cin:3d}->{cout:3d} {n}x{h}x{w} {nm:7s} rel err {err:.3e}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"fwd   {geom:7s} {cin:3d}->{cout:3d} {nm:7s} FAILED {e}", flush=True)
        xr = x.clone().requires_grad_(True)
        y = ref_conv(geom, xr, wt, b)
        dy = torch.randn(y.shape, generator=g)
        y.backward(dy)
        for math, nm in ((ops.MATH_FP32, "fp32"), (ops.MATH_TF32X3, "tf32x3")):
            try:
                dx = ops.conv_dgrad(geo, dy.cuda(), wt.cuda(), (h, w), math=math)
                torch.cuda.synchronize()
                err = float((dx.cpu() - xr.grad).abs().max()) / max(1.0, float(xr.grad.abs().max()))
                print(f"dgrad {geom:7s} {cin:3d}->{cout:3d} {n}x{h}x{w} {nm:7s} rel err {err:.3e}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"dgrad {geom:7s} {cin:3d}->{cout:3d} {nm:7s} FAILED {e}", flush=True)
    if quick:
        return
    timing()


def timing():
    # timing at the bench shapes (batch 64, ROBO-UNet 160x120 layer table)
    layers = [("k3s1d1", 3, 8, 120, 160), ("k3s2", 8, 16, 120, 160), ("k3s1d1", 16, 16, 60, 80),
              ("k3s2", 16, 32, 60, 80), ("k3s1d1", 32, 32, 30, 40), ("k3s2", 32, 64, 30, 40),
              ("k3s1d1", 64, 64, 15, 20), ("k3s1d1", 64, 128, 15, 20), ("k3s1d1", 128, 128, 15, 20),
              ("k3s1d1", 128, 64, 15, 20), ("convT", 64, 32, 15, 20), ("convT", 32, 16, 30, 40),
              ("convT", 16, 8, 60, 80), ("k1", 8, 5, 120, 160)]
    B = int(__import__("os").environ.get("PROBE_B", "64"))
    only = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--layer=")]
    if only:
        layers = [layers[i] for i in only]
    for geom, cin, cout, h, w in layers:
        k, s, p, d, tr = GEOMS[geom]
        geo = ops.ConvGeom(cin, cout, k, s, p, d, tr)
        x = torch.randn(B, cin, h, w, device="cuda")
        wt = torch.randn((cin, cout, 3, 3) if tr else (cout, cin, k, k), device="cuda") / (cin * k * k) ** 0.5
        ho, wo = geo.out_hw(h, w)
        y = torch.empty(B, cout, ho, wo, device="cuda")
        dy = torch.randn(B, cout, ho, wo, device="cuda")
        dx = torch.empty_like(x)
        px = h * w if tr else ho * wo
        fl = 2.0 * cin * cout * k * k * px * B
        row = f"{geom:7s} {cin:3d}->{cout:3d} {h:3d}x{w:3d}:"
        for math, nm in ((ops.MATH_FP32, "fp32"), (ops.MATH_TF32X3, "tf32x3")):
            try:
                tf = timeit(lambda: ops.conv_fwd(geo, x, wt, None, epilogue=ops.EPI_RELU, math=math, out=y))
                tb = timeit(lambda: ops.conv_dgrad(geo, dy, wt, (h, w), math=math, out=dx))
                dw = torch.zeros_like(wt)
                tw = timeit(lambda: ops.conv_wgrad(geo, x, dy, dw=dw, math=math))
                row += f"  {nm}: fwd {tf:6.1f} dgrad {tb:6.1f} wgrad {tw:6.1f} us"
            except Exception as e:  # noqa: BLE001
                row += f"  {nm}: FAILED {e}"
        print(row, flush=True)


if __name__ == "__main__":
    main()
