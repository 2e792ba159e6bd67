#!/usr/bin/env python
"""In-graph kernel durations of ONE replayed training step (torch.profiler / CUPTI): per-kernel totals, the sum, the
step's wall time on the device and the idle gaps of the main stream.  python tools/step_timeline.py [batch]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import synth
from robocupvision_b200.model import ROBO_UNet
from robocupvision_b200.train import TrainStep
from torch.profiler import ProfilerActivity, profile
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(12345678)
m = ROBO_UNet().cuda()
ts = TrainStep(m, synth.CLASS_WEIGHTS, lr=1e-3, l1_decay=1e-6, use_graph=True)
x = synth.images(B, 3, 120, 160, seed=3).cuda(); y = synth.labels_random(B, 120, 160, seed=4).cuda()
for _ in range(10):
    ts.step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        ts.step(x, y)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type is not None and "Memcpy" not in e.name and "Memset" not in e.name and e.device_time > 0]
ev.sort(key=lambda e: e.time_range.start)
# keep the last replay: split by the pack kernel that opens a step
starts = [i for i, e in enumerate(ev) if "pack_multi" in e.name]
seg = ev[starts[-1]:] if starts else ev
t0 = seg[0].time_range.start; t1 = max(e.time_range.end for e in seg)
import collections, re
agg = collections.OrderedDict()
for e in seg:
    k = re.sub(r"\(.*", "", e.name.replace("(anonymous namespace)::", "").replace("void ", ""))
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"{len(seg)} kernels, sum of durations {tot:.1f} us, first start -> last end {t1 - t0:.1f} us")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k[:60]:60s} {n:3d} {t:8.1f} us {100 * t / tot:5.1f}%  avg {t / n:6.1f}")
if "--trace" in sys.argv:
    # the step in start order: offset from the step's first kernel, duration, stream, idle time of that stream before it
    last_end = {}
    print("\n  start_us   dur_us  stream  idle_before_us  kernel")
    for e in seg:
        st = getattr(e, "stream", None)
        if st is None:
            st = getattr(e, "device_resource_id", -1)
        s, d = e.time_range.start - t0, e.device_time
        idle = s - last_end.get(st, s)
        last_end[st] = max(last_end.get(st, 0), e.time_range.end - t0)
        k = re.sub(r"\(.*", "", e.name.replace("(anonymous namespace)::", "").replace("void ", ""))
        print(f"  {s:8.1f} {d:8.1f}  {st!s:>6}  {idle:8.1f}        {k[:70]}")
