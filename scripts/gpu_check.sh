#!/usr/bin/env bash
# gpurun -- 'bash scripts/gpu_check.sh <tag>': full GPU parity suite, then the default bench line (with secondary workloads and CPU baselines)
set -u
TAG="${1:-chk}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --durations=5 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"
tail -8 "$OUT/pytest_gpu.log"
timeout 900 python bench.py --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
cat "$OUT/bench.json"; tail -3 "$OUT/bench.err"
