#!/bin/bash
# A/B pairs of the experiment queue in DESIGN.md section 7 (one GPU-box visit, ~2 min):
#   /usr/local/graft/bin/gpurun --timeout 300 -- 'bash tools/ab_queue.sh'
# Every line is `bench.py --steps 40 --warmup 8 --no-cpu-baseline` (training workload), value / ms per step / e2e.
set -u
mkdir -p gpurun_out
run() {  # label, env assignments...
  local label=$1; shift
  env "$@" timeout 60 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2> gpurun_out/ab_$label.err \
    | tee gpurun_out/ab_$label.json \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']))"
}
# the pending GPU tests first (first run of --v2 / FCN / channel-pruned checkpoint): XPASS expected
timeout 120 python -m pytest tests/test_gpu_zz_pending.py -q -rxX 2>&1 | tail -25
# experimental kernels: parity first (a hang here must not take the A/B lines with it: bounded)
RCV_TEST_EXPERIMENTAL=1 timeout 60 python -m pytest tests/test_gpu_zz_pending.py -q -k bn_bwd_fused 2>&1 | tail -5
run base        RCV_NOOP=1
run bn_fused    RCV_B200_BN_BWD_FUSED=1
run pdl_off     RCV_PDL=0
run bncap64     RCV_UMMA_BNCAP=64
run wgrad_nl    RCV_B200_WGRAD_ON_LOAD=1
run bn_onload0  RCV_B200_BN_ON_LOAD=0
run kb128_32    RCV_UMMA_KB128=32
run base2       RCV_NOOP=1
# descriptor probe for 16-channel (64-byte-row) halo staging, experiment 6
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I robocupvision_b200/csrc -I include -o /tmp/umma_shift_probe64 \
     tools/umma_shift_probe64.cu && timeout 60 /tmp/umma_shift_probe64 | tee gpurun_out/umma_shift_probe64.log
