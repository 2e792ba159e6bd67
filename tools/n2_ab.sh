#!/bin/bash
# 2-GPU A/B of the gradient exchange: NCCL buckets vs rcv_peer_allreduce (with and without programmatic dependent
# launch).  Run under `gpurun --gpus 2`; results in gpurun_out/n2_*.json
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 200 --warmup 20 --no-extras --no-cpu-baseline \
    > gpurun_out/n2_$tag.json 2> gpurun_out/n2_$tag.err
  echo "$tag rc=$?"
  python - <<P
import json
try:
    d = json.loads([l for l in open("gpurun_out/n2_$tag.json") if l.startswith("{")][-1])
    print("$tag", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["config"].get("gradient_exchange", "")[:20],
          "pdl", d["config"]["programmatic_dependent_launch"], "dp_check", d.get("dp_check", {}).get("ok"), d.get("dp_check", {}).get("max_weight_err"))
except Exception as e:
    print("$tag parse failed", e)
P
}
run nccl RCV_B200_DP_REDUCE=nccl
run peer RCV_B200_DP_REDUCE=peer
run peer_pdl RCV_B200_DP_REDUCE=peer RCV_PDL=1
run nccl2 RCV_B200_DP_REDUCE=nccl
run peer2 RCV_B200_DP_REDUCE=peer
run peer_pdl2 RCV_B200_DP_REDUCE=peer RCV_PDL=1
tail -5 gpurun_out/n2_peer.err
