#!/usr/bin/env bash
set -u
TAG="${1:-w}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for w in water co2; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"; echo "$w rc=$?"
  cut -c1-1600 "$OUT/bench_$w.json"; tail -3 "$OUT/bench_$w.err"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_general -s 1 -c 1 -o "$OUT/general_full" -f \
    python bench.py --workload water --steps 2 --warmup 1 --n-per-dim 60 --no-cpu-baseline --e2e-steps 1 > "$OUT/ncu_full.log" 2>&1
