#!/usr/bin/env bash
set -u
TAG="${1:-k}"; W="${2:-water}"; N="${3:-60}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_multi -s 1 -c 1 -o "$OUT/multi_full_$W" -f \
    python bench.py --workload $W --steps 2 --warmup 1 --n-per-dim $N --no-cpu-baseline --e2e-steps 1 > "$OUT/ncu_full.log" 2>&1
tail -2 "$OUT/ncu_full.log" | cut -c1-300
