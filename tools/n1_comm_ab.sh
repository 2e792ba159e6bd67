#!/bin/bash
# 1-GPU cost of the data-parallel schedule itself: plain step vs the bucketed comm-stream schedule (no exchange) vs
# the same with rcv_peer_allreduce on one rank (flags and barriers run, nothing crosses NVLink)
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/n1_$tag.json 2> gpurun_out/n1_$tag.err
  python - <<P
import json
try:
    d = json.loads([l for l in open("gpurun_out/n1_$tag.json") if l.startswith("{")][-1])
    print("$tag", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
except Exception as e:
    print("$tag parse failed", e)
P
}
for i in 1 2; do
run plain$i A=1
run sched$i RCV_B200_FORCE_COMM_PATH=1
run peer$i RCV_B200_FORCE_COMM_PATH=1 RCV_B200_DP_REDUCE=peer
done
tail -3 gpurun_out/n1_peer1.err
