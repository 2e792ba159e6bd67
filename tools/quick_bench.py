"""Quick device timing of the main paths (development aid, not the contract bench)."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch
import synth
from robocupvision_b200 import ops
from robocupvision_b200.model import ROBO_UNet, PB_FCN
from robocupvision_b200.train import TrainStep, EvalStep


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


res = {}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(12345678)
m = ROBO_UNet().cuda()
x = synth.images(B, 3, 120, 160).cuda(); y = synth.labels_random(B, 120, 160).cuda()
ts = TrainStep(m, synth.CLASS_WEIGHTS, use_graph=False)
t = timeit(lambda: ts.step(x, y), 2, 5); res["train_eager_ms"] = t
print("train eager ms", t, "kernels/step", ts.kernels_per_step, flush=True)
ts = TrainStep(m, synth.CLASS_WEIGHTS, use_graph=True)
t = timeit(lambda: ts.step(x, y), 3, 20); res["train_graph_ms"] = t
print("train graph ms", t, "fps", B / t * 1e3, flush=True)
m.eval()
with torch.no_grad():
    t = timeit(lambda: m(x), 3, 20); res["infer_eager_ms"] = t
    print("infer eager ms", t, "fps", B / t * 1e3, flush=True)
    x1 = x[:1].contiguous()
    t = timeit(lambda: m(x1), 3, 20); res["infer_b1_eager_ms"] = t
    print("infer b1 eager ms", t, flush=True)
# per-layer forward timing (eval)
plan = m._get_plan()
with torch.no_grad():
    acts = [x]
    for ti, nd in enumerate(plan.nodes):
        src = acts[nd.src]
        if nd.kind == "pool":
            acts.append(ops.maxpool2x2_fwd(src)[0]); continue
        g = nd.geom
        w = nd.conv.weight.detach(); b = nd.conv.bias.detach() if nd.conv.bias is not None else None
        f = lambda: ops.conv_fwd(g, src, w, b, epilogue=ops.EPI_RELU)
        t = timeit(f, 2, 10)
        yv = f(); acts.append(yv)
        ho, wo = yv.shape[2:]
        fl = 2 * g.cin * g.cout * g.k * g.k * (src.shape[2] * src.shape[3] if g.transposed else ho * wo) * B
        dy = torch.randn_like(yv)
        td = timeit(lambda: ops.conv_dgrad(g, dy, w, src.shape[2:]), 2, 10)
        tw = timeit(lambda: ops.conv_wgrad(g, src, dy), 2, 10)
        print(f"node {ti:2d} {g.cin:3d}->{g.cout:3d} k{g.k}s{g.stride}d{g.dil}{'T' if g.transposed else ' '} {src.shape[2]}x{src.shape[3]}: "
              f"fwd {t*1e3:7.1f}us {fl/t/1e9:7.2f} TF/s | dgrad {td*1e3:7.1f}us {fl/td/1e9:6.2f} | wgrad {tw*1e3:7.1f}us {fl/tw/1e9:6.2f}", flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(res, open("gpurun_out/quick_bench.json", "w"))
