#!/usr/bin/env bash
# gpurun -- 'bash scripts/gpu_multi_pd2.sh <tag>': multi-site parity, water / CO2 with one vs two neighbours prefetched, and an ncu capture of the water kernel
set -u
TAG="${1:-pd2}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q -x -k "site_signature" > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"; tail -3 "$OUT/pytest.log"
for wl in water co2; do for pd in 0 1; do
  LS1B200_MULTI_PD2=$pd timeout 300 python bench.py --workload $wl --steps 5 --warmup 2 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_${wl}_pd$pd.json" 2> "$OUT/bench_${wl}_pd$pd.err"
  python - "$OUT/bench_${wl}_pd$pd.json" "$wl pd2=$pd" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-22s MMUPS %8.1f  force ms %.3f" % (sys.argv[2], d["value"], d["phases_ms_per_step"]["force"]))
except Exception as e:
    print(sys.argv[2], "no bench line:", e)
PY
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_multi_tiled -s 2 -c 1 -o "$OUT/multi_full_water" -f \
    python bench.py --workload water --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 1 > "$OUT/ncu_water.log" 2>&1
ls -la "$OUT"
