#!/usr/bin/env python
"""bench.py -- the hot path's throughput on B200, one JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload train|infer|infer_pbfcn|train_unet|train_lp|infer_vga] [--batch B]

Default workload (N=1): BASELINE.json configs[1] -- ROBO-UNet 160x120 TRAINING, batch 64 per
GPU, synthetic images/labels, 5 classes; a "step" is one full train step (forward, weighted
CE, + L1 term, backward, Adam) over one batch of synthetic frames.  value = frames/s over all
ranks with inputs resident in HBM; e2e = the same through robocupvision_b200.train.TrainStep
with pinned-host inputs copied H2D and the loss read back D2H every step.

--impl reference times the reference's CPU implementation of the same step on the box's host
cores: the oracle port (oracle/ref_train.py drives the same ATen CPU kernels the reference's
model.py dispatches to; the reference itself is Python under /root/reference and cannot travel).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import synth  # noqa: E402

# algorithmic bytes / flops per frame (fp32 activations written once
# and read once per consumer; conv flops = 2*MAC), per workload.
WORKLOADS = {
    #                 net / ctor                                   shape           MFLOP/frame  MB/frame  MB/step
    "train":       dict(net="ROBO_UNet", kw={}, cin=3, h=120, w=160, train=True, mflop=1489.3, mb=25.805, mb_step=19.30,
                        name="ROBO-UNet 160x120 training, batch 64/GPU, 5 classes (BASELINE configs[1])", batch=64),
    "infer":       dict(net="ROBO_UNet", kw={}, cin=3, h=120, w=160, train=False, mflop=496.4, mb=8.602, mb_step=2.758,
                        name="ROBO-UNet 160x120 inference", batch=256),
    "infer_pbfcn": dict(net="PB_FCN", kw=dict(noScale=False), cin=3, h=120, w=160, train=False, mflop=540.7, mb=9.370,
                        mb_step=2.746, name="PB_FCN 160x120 inference (BASELINE configs[2] shape)", batch=256),
    "infer_vga":   dict(net="PB_FCN", kw=dict(noScale=True), cin=3, h=480, w=640, train=False, mflop=4005.9, mb=132.096,
                        mb_step=2.857, name="PB_FCN 640x480 inference (BASELINE configs[3])", batch=8),
    "train_unet":  dict(net="ROBO_UNet", kw=dict(pool=True, levels=3, bellySize=0), cin=3, h=120, w=160, train=True,
                        mflop=494.0, mb=25.344, mb_step=2.76, name="U-Net (--UNet) 160x120 training", batch=64),
    "train_lp":    dict(net="LabelProp", kw={}, cin=8, h=120, w=160, train=True, mflop=357.6, mb=23.040, mb_step=2.58,
                        name="LabelProp two-frame training (16 samples = 8 frame pairs)", batch=16),
}
ENGINE_NAMES = {0: "igemm (fp32 FFMA)", 1: "direct_conv (fp32 FFMA)", 2: "umma_halo / umma_igemm (tcgen05 3xTF32, halo-staged A operand for stride-1 3x3)",
                3: "narrow_conv (TMA halo staging + FFMA2)"}  # rcv_engine
METRIC = "robo_unet_160x120_train_frames_per_sec"
UNIT = "frames/s"


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        p = json.loads(f.read_text())
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"],
                    tensor_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self) -> int:
        """Rows read so far: stop(first=mark()) keeps the samples taken after this point."""
        return len(self.rows)

    def stop(self, first: int = 0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[first:]:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_model(wl, device):
    from robocupvision_b200 import model as M
    torch.manual_seed(12345678)  # train.py:332
    if wl["net"] == "ROBO_UNet":
        m = M.ROBO_UNet(**wl["kw"])
    elif wl["net"] == "PB_FCN":
        m = M.PB_FCN(32, 5, 1, wl["kw"]["noScale"], 0)
    else:
        m = M.LabelProp(5, 32, 0)
    return m.to(device)


def oracle_forward(wl):
    from oracle import ref_model as R
    if wl["net"] == "ROBO_UNet":
        kw = wl["kw"]
        okw = dict(pool=kw.get("pool", False), levels=kw.get("levels", 2), belly_size=kw.get("bellySize", 5))
        return lambda sd, x, training: R.robo_unet_forward(sd, x, training=training, **okw)
    if wl["net"] == "PB_FCN":
        ns = wl["kw"]["noScale"]
        return lambda sd, x, training: R.pb_fcn_forward(sd, x, ns, training=training)
    return lambda sd, x, training: R.labelprop_forward(sd, x, training=training)


def class_weights(wl):
    return synth.LP_CLASS_WEIGHTS if wl["net"] == "LabelProp" else synth.CLASS_WEIGHTS


def cpu_reference(wl, batch, steps, warmup):
    """The reference's CPU path for this workload via the oracle port, all host threads."""
    from oracle.ref_train import OracleTrainer
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(12345678)
    from robocupvision_b200 import model as M  # module tree only (CPU init = the reference's init)
    if wl["net"] == "ROBO_UNet":
        m = M.ROBO_UNet(**wl["kw"])
    elif wl["net"] == "PB_FCN":
        m = M.PB_FCN(32, 5, 1, wl["kw"]["noScale"], 0)
    else:
        m = M.LabelProp(5, 32, 0)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    fwd = oracle_forward(wl)
    x = synth.images(batch, wl["cin"], wl["h"], wl["w"])
    y = synth.labels_random(batch, wl["h"], wl["w"])
    if wl["train"]:
        tr = OracleTrainer(sd, fwd, class_weights(wl), lr=1e-3, l1_decay=1e-6)
        fn = lambda: tr.step(x, y)  # noqa: E731
    else:
        def fn():
            with torch.no_grad():
                return fwd(sd, x, False)
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt * 1e3, threads


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout carries exactly one JSON line: everything else that writes to file descriptor 1 (NCCL's version
    # banner at NCCL_DEBUG=VERSION/WARN, library chatter) is sent to stderr for the whole run
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=12)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    batch = args.batch or wl["batch"]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric = METRIC if args.workload == "train" else f"{args.workload}_frames_per_sec"

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, args.cpu_steps))
        fps, ms, threads = cpu_reference(wl, batch, steps, min(args.warmup, 2))
        sample = f"{steps} steps of batch {batch} after {min(args.warmup, 2)} warm-up (oracle port, torch {torch.__version__} CPU)"
        line = {"impl": "reference", "metric": metric, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 2), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "batch_per_step": batch, "host_threads": threads},
                "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the product path has no CPU fallback")
    from robocupvision_b200 import _lib, ops
    from robocupvision_b200.train import EvalStep, TrainStep
    _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    model = build_model(wl, dev)
    cw = class_weights(wl)
    nbatches = 8  # rotate distinct input batches; the per-step activation working set is >> L2 anyway
    xs_host = [synth.images(batch, wl["cin"], wl["h"], wl["w"], seed=1234 + 17 * rank + i).pin_memory() for i in range(nbatches)]
    ys_host = [synth.labels_random(batch, wl["h"], wl["w"], seed=4321 + 17 * rank + i).pin_memory() for i in range(nbatches)]
    xs = [t.to(dev) for t in xs_host]
    ys = [t.to(dev) for t in ys_host]

    if wl["train"]:
        ts = TrainStep(model, cw, lr=1e-3, l1_decay=1e-6, use_graph=not args.no_graph)
        ts.broadcast_state(0)
        step = lambda i: ts.step(xs[i % nbatches], ys[i % nbatches])  # noqa: E731
        result = lambda: ts.loss_sums  # noqa: E731
    else:
        ev = EvalStep(model, cw, use_graph=not args.no_graph)
        out = {}

        def step(i):
            out["r"] = ev(xs[i % nbatches], ys[i % nbatches])
        result = lambda: out["r"]["loss"]  # noqa: E731

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # nvidia-smi needs a few hundred ms before its first row: start it ahead of the warm-up, keep the rows that
    # arrive from the start of the timed region on
    sampler = ClockSampler(local_rank)
    if rank == 0:  # one poller per job (rank 0's GPU): NVML queries are not free
        sampler.start()
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    k0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    row0 = sampler.mark()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_total = float(t)
    k1 = ops.launch_count()
    # a timed region shorter than a few sampling periods: keep the same load running (untimed) until three samples
    # have been taken under it, and say so
    extra = 0
    per_round = max(8, int(200.0 / max(ms_total / args.steps, 1e-3)))  # ~0.2 s of steps
    for _ in range(8):
        need = 1 if (sampler.proc and sampler.mark() - row0 < 3) else 0
        if world > 1:  # every rank runs the same number of steps (the step holds a collective)
            t = torch.tensor([need], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            need = int(t)
        if not need:
            break
        for i in range(per_round):
            step(i)
        extra += per_round
        barrier()
    clocks = sampler.stop(first=row0)
    clocks["window"] = "timed region" if extra == 0 else f"timed region + {extra} untimed steps of the same load"
    ms_step = ms_total / args.steps
    value = batch * world / (ms_step * 1e-3)
    if wl["train"]:
        launches = ts.kernels_per_step * args.steps
    elif ev.use_graph:
        launches = ev.kernels_per_call * args.steps
    else:
        launches = k1 - k0

    # ---- e2e: pinned host inputs -> H2D -> step -> D2H of the loss, every step ----------------
    # The public pipelined API (TrainStep.step_async / EvalStep.run_async): step i's inputs are
    # copied H2D on a copy stream while step i-1 still computes, and every step's loss is copied
    # D2H behind it and read by the host one step later.  All copies happen inside the timed
    # region, once per step.
    def e2e_loop(n):
        prev = None
        for i in range(n):
            if wl["train"]:
                h = ts.step_async(xs_host[i % nbatches], ys_host[i % nbatches])
                if prev is not None:
                    prev.wait()
            else:
                h = ev.run_async(xs_host[i % nbatches], ys_host[i % nbatches])
                if prev is not None:
                    EvalStep.wait_host(prev)
            prev = h
        return prev.wait() if wl["train"] else EvalStep.wait_host(prev)
    e2e_loop(4)
    barrier()
    e2e_steps = args.steps
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(t)
    e2e_value = batch * world / (e2e_ms * 1e-3)
    h2d = xs_host[0].numel() * 4 + ys_host[0].numel() * 8
    d2h = 32 if wl["train"] else 16

    # ---- roofline of the dominant kernel, measured live with CUDA events -----------------------
    pk = peaks()
    roof = dominant_kernel_roofline(model, wl, batch, dev, pk)
    roof_step = {"bound": "hbm", "achieved": (wl["mb"] * batch + wl["mb_step"]) * 1e6 / (ms_step * 1e-3) / 1e9,
                 "peak": pk["hbm"], "unit": "GB/s"}
    roof_step["frac"] = roof_step["achieved"] / roof_step["peak"]

    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": batch, "global_batch": batch * world,
                       "input": f"{wl['cin']}x{wl['h']}x{wl['w']}", "parallelism": f"dp{world}",
                       "cuda_graph": bool(not args.no_graph),
                       "programmatic_dependent_launch": bool(_lib.load().rcv_get_pdl()),
                       "l2": "8 rotating input batches; per-step activation working set >> 126 MB L2",
                       "peaks": pk["source"]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roof, "roofline_whole_step": roof_step,
            "roofline_hbm_layer": narrow_layer_roofline(model, wl, batch, dev, pk)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, ms, threads = cpu_reference(wl, batch, args.cpu_steps, 2)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": ms,
                                "sample": f"{args.cpu_steps} steps of batch {batch} after 2 warm-up "
                                          f"(oracle port of train.py:43-74 on torch {torch.__version__} CPU)"}
    if rank == 0:
        emit(line)
    if world > 1:
        # NCCL communicators referenced by captured CUDA graphs can stall destroy_process_group()
        # at interpreter teardown (seen on 2xB200: line printed, ranks never exited).  All ranks
        # meet once more, drain their streams and leave without running the teardown.
        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def dominant_kernel_roofline(model, wl, batch, dev, pk):
    """Time the kernel that dominates the step -- the implicit-GEMM conv on the widest layer --
    with CUDA events on the launching stream, on this workload's real tensors."""
    from robocupvision_b200 import ops
    plan = model._get_plan()
    # widest conv node = most flops per launch
    best, best_fl = None, -1
    h, w = wl["h"], wl["w"]
    shapes = {0: (wl["cin"], h, w)}
    for t, nd in enumerate(plan.nodes):
        c, hh, ww = shapes[nd.src]
        if nd.kind == "pool":
            shapes[t + 1] = (c, hh // 2, ww // 2)
            continue
        ho, wo = nd.geom.out_hw(hh, ww)
        cout = nd.geom.cout * (2 if nd.skip_mode == "cat" and nd.skip >= 0 else 1)
        shapes[t + 1] = (cout, ho, wo)
        px = hh * ww if nd.geom.transposed else ho * wo
        fl = 2.0 * nd.geom.cin * nd.geom.cout * nd.geom.k ** 2 * px
        if fl > best_fl:
            best, best_fl, best_in = nd, fl, (c, hh, ww)
    g = best.geom
    x = torch.randn(batch, *best_in, device=dev)
    wt = best.conv.weight.detach()
    ho, wo = g.out_hw(best_in[1], best_in[2])
    y = torch.empty(batch, g.cout, ho, wo, device=dev)
    eng = ops.conv_engine(g, batch, best_in[1], best_in[2], ops.PACK_FWD, ops.MATH_AUTO)
    on_tc = eng == ops.ENGINE_UMMA
    wp = ops.conv_pack(g, wt, ops.PACK_FWD) if on_tc else None
    flush = torch.empty(64 * 1024 * 1024, device=dev)  # 256 MB > L2
    times = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv_fwd(g, x, wt, None, epilogue=ops.EPI_RELU, out=y, math=ops.MATH_AUTO, wpacked=wp)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    flops = best_fl * batch
    bytes_alg = 4.0 * (x.numel() + y.numel() + wt.numel())
    intensity = flops / bytes_alg
    ridge = pk["tensor_sustained"] * 1e12 / (pk["hbm"] * 1e9)
    tfl = flops / (ms * 1e-3) / 1e12
    if intensity >= ridge * 0.5:
        # the burst figure: this kernel is timed alone.  The fp32-parity mode spends three TF32 MMAs per
        # product and TF32 runs at half the bf16 rate, so its ceiling is peak/6 (also reported).
        achieved, peak, unit, bound = tfl, pk["tensor_burst"], "TFLOP/s", "tensor"
    else:
        achieved, peak, unit, bound = bytes_alg / (ms * 1e-3) / 1e9, pk["hbm"], "GB/s", "hbm"
    out = {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak, "traffic": None,
           "kernel": f"{ENGINE_NAMES.get(eng, str(eng))} conv fwd {g.cin}->{g.cout} "
                     f"k{g.k} s{g.stride} d{g.dil} @{best_in[1]}x{best_in[2]} batch {batch}",
           "us_per_launch": ms * 1e3, "flop_per_byte": intensity,
           "math": "tcgen05 kind::tf32 x3 (fp32-level accuracy), TMEM accumulators" if on_tc else "fp32 FMA (CUDA cores)",
           "achieved_tflops": tfl, "achieved_gbs": bytes_alg / (ms * 1e-3) / 1e9, "peak_is": "measured bf16 dense (burst)"}
    tfile = ROOT / "profiles" / "r1_traffic.json"
    if tfile.exists():
        key = f"conv fwd {g.cin}->{g.cout} k{g.k} s{g.stride} d{g.dil} @{best_in[1]}x{best_in[2]} batch {batch}"
        ent = json.loads(tfile.read_text()).get(key)
        if ent:
            out["traffic"] = ent["bytes"]
            out["traffic_source"] = ent["source"]
            out["algorithmic_bytes"] = bytes_alg
    if on_tc:
        out["tf32_mma_tflops"] = 3.0 * tfl
        out["frac_of_tf32x3_ceiling"] = tfl / (pk["tensor_burst"] / 6.0)
    return out


def narrow_layer_roofline(model, wl, batch, dev, pk):
    """The HBM-shaped end of the net: the first conv node (3 or 8 input channels at full resolution) on the
    narrow-layer engine (TMA halo staging + FFMA2), timed alone with the L2 flushed, against measured HBM
    bandwidth.  Algorithmic bytes = input read once + output written once + weights."""
    from robocupvision_b200 import _lib, ops
    plan = model._get_plan()
    nd = next(n for n in plan.nodes if n.kind == "conv")
    g = nd.geom
    h, w = wl["h"], wl["w"]
    x = torch.randn(batch, g.cin, h, w, device=dev)
    wt = nd.conv.weight.detach()
    ho, wo = g.out_hw(h, w)
    y = torch.empty(batch, g.cout, ho, wo, device=dev)
    eng = ops.conv_engine(g, batch, h, w, ops.PACK_FWD, ops.MATH_AUTO)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    times = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv_fwd(g, x, wt, None, epilogue=ops.EPI_RELU, out=y, math=ops.MATH_AUTO)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    bytes_alg = 4.0 * (x.numel() + y.numel() + wt.numel())
    flops = 2.0 * g.cin * g.cout * g.k ** 2 * ho * wo * batch
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    names = ENGINE_NAMES
    out = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": None,
           "kernel": f"{names.get(eng, str(eng))} conv fwd {g.cin}->{g.cout} k{g.k} s{g.stride} d{g.dil} @{h}x{w} batch {batch}",
           "us_per_launch": ms * 1e3, "algorithmic_bytes": bytes_alg, "flop_per_byte": flops / bytes_alg,
           "fp32_tflops": flops / (ms * 1e-3) / 1e12}
    tfile = ROOT / "profiles" / "r1_traffic.json"
    if tfile.exists():
        ent = json.loads(tfile.read_text()).get(
            f"conv fwd {g.cin}->{g.cout} k{g.k} s{g.stride} d{g.dil} @{h}x{w} batch {batch}")
        if ent:
            out["traffic"], out["traffic_source"] = ent["bytes"], ent["source"]
    return out


if __name__ == "__main__":
    sys.exit(main())
