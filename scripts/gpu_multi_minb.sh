#!/usr/bin/env bash
# gpurun -- 'bash scripts/gpu_multi_minb.sh <tag>': multi-site kernel, register budget (CTAs per SM the allocation is tuned for) on water 4 M / CO2 8 M
set -u
TAG="${1:-minb}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q -x -k "site_signature" > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"; tail -3 "$OUT/pytest.log"
for wl in water co2; do for mb in 0 6 7; do
  LS1B200_MULTI_MINB=$mb timeout 300 python bench.py --workload $wl --steps 5 --warmup 2 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_${wl}_mb$mb.json" 2> "$OUT/bench_${wl}_mb$mb.err"
  python - "$OUT/bench_${wl}_mb$mb.json" "$wl minb=$mb" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-22s MMUPS %8.1f  force ms %.3f" % (sys.argv[2], d["value"], d["phases_ms_per_step"]["force"]))
except Exception as e:
    print(sys.argv[2], "no bench line:", e)
PY
done; done
