#!/usr/bin/env bash
# gpurun -- 'bash scripts/gpu_v3_sweep.sh <tag>': parity of every variant of the v3 1CLJ kernel (fine bins S, lane-group alignment,
# trim on/off) + the round-1 kernel, then a short bench line per variant (force phase ms, 16 M molecules).
set -u
TAG="${1:-v3}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
run() {   # name, env...
  local name="$1"; shift
  ( export "$@"; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decomposition.py -m gpu -x -q -k "1clj or lattice or blocks_on_one or trace or border or simulate or determin" > "$OUT/pytest_$name.log" 2>&1; echo "$name pytest rc=$?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/pytest_$name.log" | tee -a "$OUT/summary.txt"
    timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"; echo "$name bench rc=$?" | tee -a "$OUT/summary.txt"
    python - "$OUT/bench_$name.json" "$name" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   %-22s MMUPS %8.1f  ms/step %.3f  phases %s" % (sys.argv[2], d["value"], d["ms_per_step"], d.get("phases_ms_per_step")))
except Exception as e:
    print("   ", sys.argv[2], "no bench line:", e)
PY
  )
}
run tiled      LS1B200_LJ_KERNEL=tiled LS1B200_FINE_S=1
run s1a1       LS1B200_FINE_S=1 LS1B200_V3_ALIGN=1
run s1a2       LS1B200_FINE_S=1 LS1B200_V3_ALIGN=2
run s1a4       LS1B200_FINE_S=1 LS1B200_V3_ALIGN=4
run s4a1       LS1B200_FINE_S=4 LS1B200_V3_ALIGN=1
run s4a2       LS1B200_FINE_S=4 LS1B200_V3_ALIGN=2
run s2a1       LS1B200_FINE_S=2 LS1B200_V3_ALIGN=1
run s2a2       LS1B200_FINE_S=2 LS1B200_V3_ALIGN=2
run s8a1       LS1B200_FINE_S=8 LS1B200_V3_ALIGN=1
run s4a1notrim LS1B200_FINE_S=4 LS1B200_V3_ALIGN=1 LS1B200_V3_TRIM=0
cat "$OUT/summary.txt"
timeout 600 python -m pytest tests -m gpu -q > "$OUT/pytest_full_default.log" 2>&1; echo "full default pytest rc=$?" | tee -a "$OUT/summary.txt"; tail -15 "$OUT/pytest_full_default.log"
