"""CPU suite: the C-ABI library loads and exports every symbol include/rcv_b200.h declares (no
compute calls without a GPU), argument validation returns rcv_status codes, and the host-side
logic (plans, legacy checkpoint loading, train-step bookkeeping, gloo data-parallel averaging)."""
import ctypes as C
import os
import re
import socket
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    h = (ROOT / "include" / "rcv_b200.h").read_text()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(rcv_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_exported_and_bound():
    from robocupvision_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in rcv_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.rcv_version() == 5


def test_validation_without_gpu():
    from robocupvision_b200 import _lib
    lib = _lib.load()
    ho, wo = C.c_int32(), C.c_int32()
    d = _lib.ConvDesc(2, 3, 120, 160, 8, 3, 2, 1, 1, 0, 0, 0)
    assert lib.rcv_conv_out_hw(C.byref(d), C.byref(ho), C.byref(wo)) == 0 and (ho.value, wo.value) == (60, 80)
    d = _lib.ConvDesc(2, 64, 15, 20, 32, 3, 2, 1, 1, 1, 0, 0)
    assert lib.rcv_conv_out_hw(C.byref(d), C.byref(ho), C.byref(wo)) == 0 and (ho.value, wo.value) == (30, 40)
    bad = _lib.ConvDesc(2, 3, 120, 160, 8, 5, 1, 2, 1, 0, 0, 0)
    assert lib.rcv_conv_out_hw(C.byref(bad), C.byref(ho), C.byref(wo)) == _lib.RCV_ERR_UNSUPPORTED
    assert b"kernel size 5" in lib.rcv_last_error()
    assert lib.rcv_conv_fwd(C.byref(d), None, None, None, None, None, None, None, None, None, None) == _lib.RCV_ERR_BAD_ARG
    assert lib.rcv_conv_packed_bytes(C.byref(d), 0) == 4 * 1 * 8 * 32 * 256  # 4 parity classes, K=64*4 -> 8 blocks, BN=32
    assert lib.rcv_ce_fwd(1, 9, 10, None, None, None, None, None, None, None, None) == _lib.RCV_ERR_BAD_ARG
    assert lib.rcv_adam_l1_step(0, None, None, None, None, None, 0.1, 0.9, 0.999, 1e-8, 1, 0.0, 1.0, None, None,
                                None, None) == _lib.RCV_ERR_BAD_ARG


def test_no_cpu_fallback():
    from robocupvision_b200.model import ROBO_UNet, CrossEntropyLoss2d
    m = ROBO_UNet()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 24, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        CrossEntropyLoss2d()(torch.zeros(1, 5, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))


def test_product_never_imports_oracle():
    for f in (ROOT / "robocupvision_b200").rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"


def test_plans():
    from robocupvision_b200.model import ROBO_UNet, PB_FCN, LabelProp, FCN
    p = ROBO_UNet()._get_plan()
    assert len(p.nodes) == 16 and len(p.params) == 62
    assert [n.skip for n in p.nodes[12:15]] == [5, 3, 1]       # Up_i + downs[-(i+2)]  (model.py:509)
    p = ROBO_UNet(pool=True, levels=3, bellySize=0)._get_plan()
    assert [n.kind for n in p.nodes].count("pool") == 3
    p = PB_FCN(32, 5, 1, True, 0)._get_plan()
    assert len(p.nodes) == 18
    assert all(not id(q) in {id(x) for x in p.params} for q in PB_FCN(32, 5, 1, True, 0).classifier.parameters())
    p = LabelProp(5, 32, 0)._get_plan()
    assert p.nodes[9].skip_mode == "partial" and p.nodes[9].skip_ch == 8
    assert len(FCN()._get_plan().nodes) == 15
    m = ROBO_UNet()
    with pytest.raises(KeyError):
        m.downPart[0]                      # reference quirk: add_module names, slices still work
    assert len(list(m.downPart[0:2].parameters())) == 12


def test_legacy_checkpoint_loading():
    from robocupvision_b200.model import PB_FCN, LabelProp, load_legacy_state_dict
    from util import load_ckpt
    raw = load_ckpt("bestModelSeg")
    assert "classifier.classifier.weight" in raw and "segmenter.classifier.weight" not in raw
    m = PB_FCN(32, 5, 1, False, 0)
    with pytest.raises(RuntimeError):
        m.load_state_dict(raw)                                   # strict load fails, as in the reference
    missing, unexpected = load_legacy_state_dict(m, raw)
    assert not missing and not unexpected
    assert torch.equal(m.segmenter.classifier.weight, raw["classifier.classifier.weight"])
    with pytest.raises(RuntimeError):
        load_legacy_state_dict(PB_FCN(32, 5, 1, True, 0), raw)   # VGA net needs conv_ext / up4
    lp = LabelProp(5, 32, 0)
    load_legacy_state_dict(lp, load_ckpt("bestModelLPFinetunedPruned"))
    z = float((lp.conv2.conv.weight == 0).float().mean())
    assert z > 0.5                                               # magnitude-pruned, dense shapes


def test_pruning_helpers():
    from robocupvision_b200.model import ROBO_UNet, pruneModelNew, count_zero_weights, getParamSize
    torch.manual_seed(0)
    m = ROBO_UNet()
    with torch.no_grad():
        masks = pruneModelNew(m.parameters(), ratio=0.2)
    big = [p for p in m.parameters() if p.dim() > 1]
    assert len(masks) == len(big) == 16
    for p, mk in zip(big, masks):
        assert mk.shape == p.shape and float(p[mk].abs().sum()) == 0.0
    assert 0.1 < count_zero_weights(m) < 0.4
    assert getParamSize(big[0]) == big[0].numel()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import synth
    from oracle import ref_model as R
    from robocupvision_b200 import dp
    from robocupvision_b200.model import ROBO_UNet
    from robocupvision_b200.train import flatten_parameters, sync_bn_buffers
    torch.manual_seed(12345678)
    m = ROBO_UNet()
    sd = R.leaf_state_dict({k: v.clone() for k, v in m.state_dict().items()})
    # rank r owns samples [r*B, (r+1)*B) of the global batch; local BN statistics (no SyncBN)
    B = 2
    x = synth.images(B * world, 3, 24, 32, seed=5)[rank * B:(rank + 1) * B]
    y = synth.labels_learnable(synth.images(B * world, 3, 24, 32, seed=5))[rank * B:(rank + 1) * B]
    loss = R.cross_entropy_2d(R.robo_unet_forward(sd, x, training=True), y, torch.tensor(synth.CLASS_WEIGHTS))
    loss.backward()
    # the PRODUCT's arena layout and bucket plan (the exact host logic TrainStep runs), fed the oracle's gradients
    arena, table = flatten_parameters(m)
    plan = m._get_plan()
    offsets = {id(p): (o, k) for p, o, k in table}
    buckets = dp.plan_buckets([[offsets[id(p)] for p in nd.params()] for nd in plan.nodes],
                              [(o, k) for _, o, k in table], arena.numel(), 3)
    flat = torch.zeros_like(arena)
    for (name, p), (_, o, k) in zip(m.named_parameters(), table):
        flat[o:o + k] = sd[name].grad.reshape(-1)
    local = flat.clone()
    dp.allreduce_buckets(flat, buckets)
    flat /= world
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world
    ok = torch.allclose(flat, expect, rtol=0, atol=1e-7)
    # BatchNorm buffers: rank-local after training on different shards; sync_bn_buffers makes the eval models equal
    osd = {k: v.detach().clone() for k, v in sd.items()}  # running stats were updated by this rank's shard
    m.load_state_dict(osd)
    before = [torch.zeros(8) for _ in range(world)]
    dist.all_gather(before, m.state_dict()["downPart.Level0.layers.Conv0.bn.running_mean"].clone())
    differ = not torch.equal(before[0], before[1])
    sync_bn_buffers(m)
    xe = synth.images(1, 3, 24, 32, seed=9)
    with torch.no_grad():
        logits = R.robo_unet_forward({k: v.clone() for k, v in m.state_dict().items()}, xe, training=False)
    lg = [torch.zeros_like(logits) for _ in range(world)]
    dist.all_gather(lg, logits.contiguous())
    same = all(torch.equal(lg[0], t) for t in lg[1:])
    if rank == 0:
        out.put(bool(ok and differ and same))
    dist.destroy_process_group()


def test_plan_buckets_partition_and_order():
    """dp.plan_buckets on every model family's real plan: the buckets partition the arena, come in completion order,
    and every parameter of a node >= first_node lies inside [start, total) (so the bucket's gradients are final once
    backward has passed first_node)."""
    from robocupvision_b200 import dp
    from robocupvision_b200.model import FCN, PB_FCN, LabelProp, ROBO_UNet
    from robocupvision_b200.train import ARENA_ALIGN
    nets = [ROBO_UNet(), ROBO_UNet(noScale=True), ROBO_UNet(pool=True, levels=3, bellySize=0), ROBO_UNet(v2=True),
            PB_FCN(32, 5, 1, False, 0), PB_FCN(32, 5, 1, True, 0), LabelProp(5, 32, 0), FCN()]
    for m in nets:
        plan = m._get_plan()
        offs, table, o = {}, [], 0
        for p in m.parameters():
            offs[id(p)] = (o, p.numel()); table.append((o, p.numel()))
            o += -(-p.numel() // ARENA_ALIGN) * ARENA_ALIGN
        node_params = [[offs[id(p)] for p in nd.params()] for nd in plan.nodes]
        for nb in (1, 2, 3, 5):
            b = dp.plan_buckets(node_params, table, o, nb)
            assert 1 <= len(b) <= nb and b[0][2] == o and b[-1][0] == 0 and b[-1][1] == 0
            for (t0, s0, e0), (t1, s1, e1) in zip(b[:-1], b[1:]):
                assert s0 == e1 and t0 > t1 and s0 < e0
            for t, start, end in b:
                for tt in range(t, len(plan.nodes)):
                    assert all(off >= start for off, _ in node_params[tt]), (type(m).__name__, nb, t, tt)
        if len(plan.nodes) > 8:
            assert len(dp.plan_buckets(node_params, table, o, 3)) == 3


def test_data_parallel_gradient_averaging_gloo():
    """world_size-2 gloo: the product's arena layout + bucket plan (train.flatten_parameters, dp.plan_buckets:
    exactly what TrainStep runs), all-reduced bucket by bucket == mean of the per-rank oracle gradients; and
    train.sync_bn_buffers makes the rank-local BatchNorm buffers (hence the eval logits) identical."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _rank_batches(rank, nb=3, B=2, H=12, W=16, C=5):
    g = torch.Generator().manual_seed(100 + rank)
    return [(torch.randint(0, C, (B, H, W), generator=g), torch.randint(0, C, (B, H, W), generator=g),
             float(torch.rand(1, generator=g))) for _ in range(nb)]


def _meter_worker(rank, world, port, out):
    import numpy as np
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ref_metrics as RM
    from robocupvision_b200.train import ValidationMeter, iou_sums
    C = 5
    meter = ValidationMeter(C, "cpu")
    for pred, tgt, loss in _rank_batches(rank):
        conf = torch.from_numpy(RM.confusion_per_image(pred.numpy(), tgt.numpy(), C))
        meter.update({"conf": conf, "iou_sum": iou_sums(conf), "correct": (pred == tgt).sum(), "argmax": pred,
                      "loss": torch.tensor(loss)})
    got = meter.summary()
    # the oracle over the union of every rank's batches (train.py:133-164)
    conf_tot, iou, imgs, correct, pixels, losses = np.zeros((C, C), np.int64), np.zeros(C), 0, 0, 0, []
    for r in range(world):
        for pred, tgt, loss in _rank_batches(r):
            cpi = RM.confusion_per_image(pred.numpy(), tgt.numpy(), C)
            conf_tot += cpi.sum(0); iou += RM.iou_sums(cpi); imgs += pred.shape[0]
            correct += int((pred == tgt).sum()); pixels += pred.numel(); losses.append(loss)
    mca, miou, score = RM.epoch_summary(conf_tot, iou, imgs)
    ok = (abs(got["mean_class_acc"] - mca) < 1e-9 and abs(got["mean_iou"] - miou) < 1e-9 and
          abs(got["score"] - score) < 1e-9 and got["images"] == imgs and
          abs(got["pixel_acc"] - 100.0 * correct / pixels) < 1e-9 and
          abs(got["loss"] - sum(losses) / len(losses)) < 1e-6 and
          bool((got["conf"].numpy() == conf_tot).all()))
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_validation_meter_reduces_over_ranks_gloo():
    """world_size-2 gloo: ValidationMeter's epoch summary (pixel accuracy, column-normalised confusion, mean class
    accuracy, mean IoU, score) over image-sharded ranks == the oracle's train.py:133-164 over all images."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_meter_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_engine_dispatch_table_without_gpu():
    """rcv_conv_engine is pure host logic: the dispatch of every ROBO-UNet layer
    at the bench batch -- narrow-layer engine where <= 16 channels sit on the narrow side, tensor cores for the
    rest, and the CUDA-core fallbacks for widths the TMA tensor maps cannot take."""
    from robocupvision_b200 import _lib, ops
    N, U, S = _lib.ENGINE_NARROW, _lib.ENGINE_UMMA, _lib.ENGINE_SIMT
    table = [  # geom, cin, cout, h, w -> (fwd, dgrad, wgrad)
        (("k3s1d1", 3, 8, 120, 160), (N, N, N)),
        (("k3s2", 8, 16, 120, 160), (N, N, N)),
        (("k3s1d1", 16, 16, 60, 80), (U, U, N)),   # persistent 16-channel tensor-core kernel; weight gradient narrow
        (("k3s2", 16, 32, 60, 80), (U, N, U)),
        (("k3s1d1", 32, 32, 30, 40), (U, U, U)),
        (("k3s1d1", 128, 128, 15, 20), (U, U, U)),
        (("convT", 64, 32, 15, 20), (U, U, U)),
        (("convT", 32, 16, 30, 40), (N, U, U)),
        (("convT", 16, 8, 60, 80), (N, N, N)),
        (("k1", 8, 5, 120, 160), (N, N, N)),
    ]
    geoms = {"k3s1d1": (3, 1, 1, 1, False), "k3s2": (3, 2, 1, 1, False), "k1": (1, 1, 0, 1, False),
             "convT": (3, 2, 1, 1, True)}
    for (geom, cin, cout, h, w), want in table:
        k, s, p, d, tr = geoms[geom]
        g = ops.ConvGeom(cin, cout, k, s, p, d, tr)
        got = tuple(ops.conv_engine(g, 64, h, w, direction, ops.MATH_AUTO) for direction in (0, 1, 2))
        assert got == want, (geom, cin, cout, got, want)
    # a width that is not a multiple of 4 cannot be a TMA row: no narrow engine
    g = ops.ConvGeom(3, 8, 3, 1, 1, 1, False)
    assert ops.conv_engine(g, 2, 9, 7, 0, ops.MATH_AUTO) != N
    # RCV_MATH_FP32 never picks the tensor cores
    g = ops.ConvGeom(128, 128, 3, 1, 1, 1, False)
    assert ops.conv_engine(g, 64, 15, 20, 0, ops.MATH_FP32) == S


def test_pb_fcn_2_is_the_unet_trunk_plus_a_classification_head():
    """model.py:416-459: PB_FCN_2 registers the ROBO_UNet default trunk under the same names plus `classifier`; its
    segmentation branch builds a plan with the same 16 conv nodes, the classification branch is refused."""
    from robocupvision_b200.model import PB_FCN_2, ROBO_UNet
    torch.manual_seed(3)
    m2, mu = PB_FCN_2(False), ROBO_UNet()
    k2, ku = list(m2.state_dict().keys()), list(mu.state_dict().keys())
    assert [k for k in k2 if not k.startswith("classifier.")] == ku
    assert [k for k in k2 if k.startswith("classifier.")] == ["classifier.layers.Class.weight",
                                                              "classifier.layers.Class.bias"]
    assert all(m2.state_dict()[k].shape == mu.state_dict()[k].shape for k in ku)
    p2, pu = m2._get_plan(), mu._get_plan()
    assert [(n.kind, n.src, n.skip, n.order) for n in p2.nodes] == [(n.kind, n.src, n.skip, n.order) for n in pu.nodes]
    with pytest.raises(NotImplementedError):
        PB_FCN_2(True)._get_plan()


def test_pdl_switch_policy(monkeypatch):
    """rcv_set_pdl / rcv_get_pdl round trip, and the loader's policy: programmatic dependent launch on for
    single-process runs, off under a multi-process launcher, RCV_PDL overrides."""
    import subprocess, sys
    from robocupvision_b200 import _lib
    lib = _lib.load()
    prev = lib.rcv_set_pdl(0)
    assert lib.rcv_get_pdl() == 0 and lib.rcv_set_pdl(1) == 0 and lib.rcv_get_pdl() == 1
    lib.rcv_set_pdl(prev)
    code = "from robocupvision_b200 import _lib; print(_lib.load().rcv_get_pdl())"
    for env, want in [({}, "1"), ({"WORLD_SIZE": "2"}, "0"), ({"WORLD_SIZE": "2", "RCV_PDL": "1"}, "1"),
                      ({"RCV_PDL": "0"}, "0")]:
        e = {k: v for k, v in os.environ.items() if k not in ("WORLD_SIZE", "RCV_PDL")}
        e.update(env)
        out = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, cwd=str(ROOT))
        assert out.stdout.strip().splitlines()[-1] == want, (env, out.stdout, out.stderr)


def test_host_side_dispatch_is_total_over_random_geometries():
    """The host-side queries of the ABI (engine choice per direction, tensor-core eligibility, packed-panel size,
    normalise-on-load eligibility) answer for ANY geometry the descriptor can express -- no crash, no out-of-range
    code -- and agree with each other: a layer that normalises on load runs on the tensor-core engine; a direction
    that uses tensor cores has a non-empty packed panel."""
    import random
    from robocupvision_b200 import ops
    from robocupvision_b200._lib import RcvError
    rnd = random.Random(1234)
    seen, refused = set(), 0
    for _ in range(3000):
        tr = rnd.random() < 0.2
        k = 3 if tr else rnd.choice([1, 3, 3, 3])
        s = 2 if tr else (1 if k == 1 else rnd.choice([1, 1, 2]))
        d = 1 if (tr or k == 1) else rnd.choice([1, 1, 2])
        p = 1 if tr else (0 if k == 1 else d)
        cin, cout = rnd.choice([1, 3, 5, 8, 16, 24, 32, 40, 64, 96, 128, 256]), rnd.choice([1, 5, 8, 16, 32, 64, 128, 256])
        g = ops.ConvGeom(cin, cout, k, s, p, d, tr)
        n, h, w = rnd.choice([1, 2, 8, 64, 256]), rnd.choice([1, 2, 5, 15, 30, 60, 120, 480]), rnd.choice([1, 3, 8, 20, 40, 80, 160, 640])
        for math in (ops.MATH_FP32, ops.MATH_AUTO):
            try:
                engines = [ops.conv_engine(g, n, h, w, direction, math) for direction in (ops.PACK_FWD, ops.PACK_DGRAD, 2)]
            except RcvError as e:  # e.g. a stride-2 layer on an odd-sized image: refused with a status and a message
                assert "-2" in str(e) and len(str(e)) > 40, str(e)
                refused += 1
                continue
            assert all(e in (ops.ENGINE_SIMT, ops.ENGINE_DIRECT, ops.ENGINE_UMMA, ops.ENGINE_NARROW) for e in engines)
            seen.update(engines)
            if math == ops.MATH_FP32:
                assert ops.ENGINE_UMMA not in engines
                assert not ops.conv_normalises_on_load(g, n, h, w, math)
                assert not ops.conv_wgrad_normalises_on_load(g, n, h, w, math)
                continue
            if ops.conv_normalises_on_load(g, n, h, w, math):
                assert engines[0] == ops.ENGINE_UMMA and k == 3 and s == 1 and not tr and cin % 32 == 0
            if ops.conv_wgrad_normalises_on_load(g, n, h, w, math):
                assert engines[2] == ops.ENGINE_UMMA and k == 3 and s == 1 and not tr and w % 4 == 0
            for direction in (ops.PACK_FWD, ops.PACK_DGRAD):
                if ops.conv_uses_tensor_cores(g, direction, math):
                    assert ops.conv_packed_bytes(g, direction) > 0
    # geometries outside the path are refused with a status and a message, never guessed at
    assert refused < 2 * 3000 * 0.5
    for bad in (ops.ConvGeom(8, 8, 1, 2, 0, 1), ops.ConvGeom(8, 8, 5, 1, 2, 1), ops.ConvGeom(8, 8, 3, 3, 1, 1)):
        with pytest.raises(RcvError):
            ops.conv_engine(bad, 2, 12, 16, ops.PACK_FWD, ops.MATH_AUTO)
    assert seen == {ops.ENGINE_SIMT, ops.ENGINE_DIRECT, ops.ENGINE_UMMA, ops.ENGINE_NARROW} or \
        seen == {ops.ENGINE_SIMT, ops.ENGINE_UMMA, ops.ENGINE_NARROW}


def test_peer_share_bounds_partition_the_range():
    """peer.share_bounds (the host statement of peer_allreduce_kernel's split): the ranks' shares are disjoint, in rank
    order, float4-aligned and cover the range, for every world size the kernel is instantiated for."""
    from robocupvision_b200.peer import share_bounds
    for count in (0, 4, 8, 36, 4 * 1001, 4 * 4096):
        for world in range(1, 9):
            end = 0
            for r in range(world):
                lo, hi = share_bounds(count, world, r)
                assert lo == end and lo <= hi and lo % 4 == 0 and hi % 4 == 0
                end = hi
            assert end == count
    with pytest.raises(ValueError):
        share_bounds(6, 2, 0)


def _peer_fail_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robocupvision_b200.peer import PeerExchange, PeerUnavailable
    try:
        PeerExchange(64, "cuda:0")
        out.put((rank, "constructed"))
    except PeerUnavailable as e:
        out.put((rank, str(e)))
    dist.barrier()  # both ranks got here: nobody was left waiting in a gather


def test_peer_exchange_unavailable_is_collective_gloo():
    """world_size-2 gloo, no GPU: the peer-memory exchange cannot be set up, and BOTH ranks learn it the same way --
    PeerUnavailable naming every rank's reason -- so TrainStep's fall-back to dist.all_reduce is taken by all ranks
    together (a rank raising on its own would leave the others waiting in the handle exchange)."""
    if torch.cuda.is_available():
        pytest.skip("the failure path needs a box without a GPU")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_fail_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(2))
    assert set(got) == {0, 1}
    for r, msg in got.items():
        assert "rank 0:" in msg and "rank 1:" in msg, (r, msg)
