#!/usr/bin/env bash
set -u
TAG="${1:-k}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_multi -s 1 -c 1 -o "$OUT/multi_full_water" -f \
    python bench.py --workload water --steps 2 --warmup 1 --n-per-dim 60 --no-cpu-baseline --e2e-steps 1 > "$OUT/ncu_w.log" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_multi -s 1 -c 1 -o "$OUT/multi_full_co2" -f \
    python bench.py --workload co2 --steps 2 --warmup 1 --n-per-dim 60 --no-cpu-baseline --e2e-steps 1 > "$OUT/ncu_c.log" 2>&1
