#!/usr/bin/env bash
# gpurun -- 'bash scripts/gpu_multi_stage.sh <tag>': parity of the multi-site kernels, then water 4 M with / without staged site records
set -u
TAG="${1:-mst}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q -x -k "site_signature or trace or multisite or water or dropin or leapfrog or golden" > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"; tail -4 "$OUT/pytest.log"
for st in 1 0; do
  LS1B200_MULTI_STAGE=$st timeout 300 python bench.py --workload water --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_water_stage$st.json" 2> "$OUT/bench_water_stage$st.err"
  python - "$OUT/bench_water_stage$st.json" "water stage=$st" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-22s MMUPS %8.1f  force ms %.3f" % (sys.argv[2], d["value"], d["phases_ms_per_step"]["force"]))
except Exception as e:
    print(sys.argv[2], "no bench line:", e)
PY
done
