#!/bin/bash
# N-GPU A/B of the gradient exchange (default = rcv_peer_allreduce vs NCCL buckets): bash tools/ngpu_ab.sh N [reps]
# under `gpurun --gpus N`; results in gpurun_out/ng<N>_*.json
N=${1:-2}; REPS=${2:-1}
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 200 --warmup 20 --no-extras --no-cpu-baseline \
    > gpurun_out/ng${N}_$tag.json 2> gpurun_out/ng${N}_$tag.err
  echo "$tag rc=$?"
  python - <<P
import json
try:
    d = json.loads([l for l in open("gpurun_out/ng${N}_$tag.json") if l.startswith("{")][-1])
    print("$tag", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["config"].get("gradient_exchange", "")[:20],
          "pdl", d["config"]["programmatic_dependent_launch"], "dp_check", d.get("dp_check", {}).get("ok"), d.get("dp_check", {}).get("max_weight_err"),
          d.get("dp_check", {}).get("reduce"))
except Exception as e:
    print("$tag parse failed", e)
P
}
for i in $(seq $REPS); do
  run peer$i A=1
  if [ -n "$FREE_RUN" ]; then run free$i RCV_B200_DP_SKIP_EXCHANGE=1; else run nccl$i RCV_B200_DP_REDUCE=nccl; fi
done
grep -il "warn\|error" gpurun_out/ng${N}_*.err | head
