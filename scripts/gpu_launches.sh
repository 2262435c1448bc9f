#!/usr/bin/env bash
set -u
TAG="${1:-l}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"; cut -c1-1500 "$OUT/bench.json"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_under_ncu.log" 2>&1
