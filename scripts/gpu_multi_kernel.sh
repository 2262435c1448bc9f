#!/usr/bin/env bash
set -u
TAG="${1:-k}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -15 "$OUT/pytest_gpu.log"
for w in water co2; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"; echo "$w rc=$?"
  python - "$OUT/bench_$w.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][0]); print(d['value'], d['phases_ms_per_step'], d['roofline']['frac'])
PY
  tail -3 "$OUT/bench_$w.err"
done
