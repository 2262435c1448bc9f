#!/usr/bin/env bash
set -u
TAG="${1:-k}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_1clj -s 2 -c 1 -o "$OUT/force_full" -f \
    python bench.py --steps 3 --warmup 3 --n-per-dim 100 --no-cpu-baseline --e2e-steps 1 > "$OUT/ncu_full.log" 2>&1
tail -1 "$OUT/ncu_full.log" | cut -c1-200
