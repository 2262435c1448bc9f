#!/usr/bin/env bash
# gpurun -- 'bash scripts/gpu_tests.sh <tag> [pytest args]': GPU parity tests only (plus a short bench line)
set -u
TAG="${1:-t}"; shift || true
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q "$@" > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"
tail -40 "$OUT/pytest_gpu.log"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_under_ncu.log" 2>&1
