#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full captures of the top kernels.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_exit=$?"
tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke_exit=$?"; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench.err; echo "bench_ref_exit=$?"
python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2>> gpurun_out/bench.err; echo "bench_exit=$?"
python bench.py --workload infer --steps 30 --warmup 5 --cpu-steps 3 > gpurun_out/bench_infer.json 2>> gpurun_out/bench.err; echo "bench_infer_exit=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu1.log 2>&1; echo "ncu1=$?"
for K in umma_halo_kernel umma_igemm_kernel narrow_conv_kernel narrow_wgrad_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:"$K" -s 20 -c 6 -f -o gpurun_out/prof_$K \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_$K.log 2>&1; echo "ncu_$K=$?"
done
ls -la gpurun_out | head -40
