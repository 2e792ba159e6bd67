#!/bin/bash
# Same-box A/B of environment switches: tools/ab.sh "NAME=VAL ..." "NAME=VAL ..." ...  (each argument one arm; "" = defaults)
# Every arm runs bench.py twice, interleaved, and the table shows frames/s and ms per step.
mkdir -p gpurun_out
for rep in 1 2; do
  i=0
  for arm in "$@"; do
    i=$((i+1))
    env $arm python bench.py --no-cpu-baseline --no-extras ${BENCH_ARGS:-} > gpurun_out/ab_${i}_${rep}.json 2>> gpurun_out/ab.err
  done
done
python - "$@" <<'PY'
import json, sys
arms = sys.argv[1:]
for i, arm in enumerate(arms, 1):
    v = []
    for rep in (1, 2):
        try:
            d = json.load(open(f"gpurun_out/ab_{i}_{rep}.json")); v.append((d["value"], d["ms_per_step"], d["roofline"]["us_per_launch"]))
        except Exception as e:
            v.append((0, 0, 0))
    print(f"{arm or '(defaults)':40s} " + "  ".join(f"{a:9.0f} fps {b:.4f} ms (top kernel {c:.1f} us)" for a, b, c in v))
PY
