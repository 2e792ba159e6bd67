#!/bin/bash
# Round-end visit on one GPU: parity suite, smoke, the default bench line, the reference arm, the ncu launch list of
# a short eager bench (after the same command has run without ncu).  Outputs under gpurun_out/final_*.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest_exit=$?"; tail -1 gpurun_out/final_pytest.log
python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke_exit=$?"; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench_exit=$?"
python bench.py --impl reference --steps 12 --warmup 3 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_ref.err; echo "ref_exit=$?"
python bench.py --math bf16 --no-extras --no-cpu-baseline > gpurun_out/final_bench_bf16.json 2>/dev/null; echo "bf16_exit=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph"
$B > gpurun_out/final_plain.log 2>&1; echo "plain_exit=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/final_launches.csv $B > gpurun_out/final_ncu.log 2>&1; echo "ncu_exit=$?"
python - <<'P'
import json
for f in ("final_bench", "final_bench_reference", "final_bench_bf16"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "parse failed", e)
P
