#!/usr/bin/env bash
set -u
TAG="${1:-lazy}"
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q -x > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
tail -4 "$OUT/pytest.log"
for f in 0 1; do
  LS1B200_LAZY=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --e2e-steps 3 > "$OUT/bench_l$f.json" 2> "$OUT/bench_l$f.err"; echo "bench lazy=$f rc=$?" | tee -a "$OUT/summary.txt"
  python - "$OUT/bench_l$f.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(d["value"], d["ms_per_step"], d["phases_ms_per_step"], d.get("trace"), d["e2e"]["value"])
PY
done
