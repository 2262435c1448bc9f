#!/usr/bin/env bash
# prefilter kernel (force_1clj_pf.cu): parity of the whole GPU suite with the kernel selected + exactness check, then the bench
set -u
TAG="${1:-pf}"
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
LS1B200_LJ_KERNEL=pf LS1B200_PF_CHECK=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_decomposition.py -m gpu -q -x > "$OUT/pytest_pf.log" 2>&1; echo "pytest pf rc=$?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest_pf.log"
for k in tiled pf; do
  LS1B200_LJ_KERNEL=$k timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --e2e-steps 1 > "$OUT/bench_$k.json" 2> "$OUT/bench_$k.err"; echo "bench $k rc=$?" | tee -a "$OUT/summary.txt"
  python - "$OUT/bench_$k.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(d["value"], d["ms_per_step"], d["phases_ms_per_step"], d.get("trace"))
PY
done
