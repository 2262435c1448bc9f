#!/usr/bin/env bash
# sweep of the variant builds of the library (scripts/build_variant.py): one parity test + bench force time each
set -u
TAG="${1:-pfsweep}"; shift
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for v in "$@"; do
  lib=ls1_mardyn_b200/variants/libls1b200_$v.so
  K=${KERNEL:-pf}
  LS1B200_LIB=$PWD/$lib LS1B200_LJ_KERNEL=$K LS1B200_PF_CHECK=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lattice_vs_oracle or one_million or golden" > "$OUT/pytest_$v.log" 2>&1; echo "$v pytest rc=$? $(tail -1 $OUT/pytest_$v.log)" | tee -a "$OUT/summary.txt"
  LS1B200_LIB=$PWD/$lib LS1B200_LJ_KERNEL=$K timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --e2e-steps 1 > "$OUT/bench_$v.json" 2> "$OUT/bench_$v.err"
  python - "$OUT/bench_$v.json" "$v" <<'PY' | tee -a "$OUT/summary.txt"
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2], d["value"], d["ms_per_step"], d["phases_ms_per_step"]["force"], d["trace"]["upot_last_per_molecule"])
PY
done
