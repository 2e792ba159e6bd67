#!/bin/bash
# One GPU-box visit: parity tests, smoke, ncu full captures of the top kernels (each after its own command has run
# without ncu).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_exit=$?"
tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke_exit=$?"; tail -1 gpurun_out/smoke.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph"
$B > gpurun_out/plain.log 2>&1; echo "plain_exit=$?"
for K in ${KERNELS:-umma_halo_kernel bn_bwd_kernel umma_wgrad_kernel narrow_conv_kernel}; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$K" -s ${SKIP:-24} -c ${COUNT:-8} -f \
      -o gpurun_out/prof_$K $B > gpurun_out/ncu_$K.log 2>&1; echo "ncu_$K=$?"
done
ls -la gpurun_out/*.ncu-rep
