#!/usr/bin/env bash
# gpurun -- 'bash scripts/gpu_multi_sweep.sh <tag>': multi-site kernel, resident CTAs per SM / L1 carve-out sweep on water 4 M and CO2 8 M
set -u
TAG="${1:-ms}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for wl in water co2; do
  for m in 0 2 3 4 5 6; do
    LS1B200_MULTI_MAXCTAS=$m timeout 300 python bench.py --workload $wl --steps 5 --warmup 2 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_${wl}_$m.json" 2> "$OUT/bench_${wl}_$m.err"
    python - "$OUT/bench_${wl}_$m.json" "$wl maxctas=$m" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-22s MMUPS %8.1f  force ms %.3f" % (sys.argv[2], d["value"], d["phases_ms_per_step"]["force"]))
except Exception as e:
    print(sys.argv[2], "no bench line:", e)
PY
  done
done
