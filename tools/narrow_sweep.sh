#!/bin/bash
# Sweep the narrow-layer engine's tile knobs on the bench layers (not a test).
for cfg in "160 56" "128 56" "192 56" "256 56" "160 40" "160 96" "256 96"; do set -- $cfg; echo "FWD threads=$1 smem=$2"; for L in 0 1 2 11 12 13; do RCV_NARROW_THREADS=$1 RCV_NARROW_SMEM_KB=$2 python tools/umma_probe.py --layer=$L 2>&1 | sed -E 's/  tf32x3.*//'; done; done
