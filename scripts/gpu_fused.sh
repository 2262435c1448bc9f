#!/usr/bin/env bash
# fused step (k_pre_key + kinetic sum in the force epilogue): whole GPU suite, then bench with and without
set -u
TAG="${1:-fused}"
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q -x > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest.log"
for f in 0 1; do
  LS1B200_FUSED=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --e2e-steps 1 > "$OUT/bench_f$f.json" 2> "$OUT/bench_f$f.err"; echo "bench fused=$f rc=$?" | tee -a "$OUT/summary.txt"
  python - "$OUT/bench_f$f.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(d["value"], d["ms_per_step"], d["phases_ms_per_step"], d.get("trace"))
PY
done
