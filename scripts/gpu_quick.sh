#!/usr/bin/env bash
# gpurun -- 'bash scripts/gpu_quick.sh <tag> [bench args]': parity tests with a tight timeout (new kernels), then a short bench line
set -u
TAG="${1:-q}"; shift || true
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 400 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"
tail -25 "$OUT/pytest_gpu.log"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
cut -c1-1400 "$OUT/bench.json"; tail -5 "$OUT/bench.err"
