#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full capture of the top kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_exit=$?"
tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench_exit=$?"
python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "bench_ref_exit=$?"
python bench.py --workload infer --steps 30 --warmup 5 --cpu-steps 3 > gpurun_out/bench_infer.json 2>> gpurun_out/bench.err; echo "bench_infer_exit=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu1.log 2>&1; echo "ncu1=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${RCV_NCU_KERNEL:-umma_igemm_kernel}" -s 20 -c 4 -f -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu2.log 2>&1; echo "ncu2=$?"
ls -la gpurun_out | head -30
