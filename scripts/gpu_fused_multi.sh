#!/usr/bin/env bash
# gpurun --gpus N -- 'bash scripts/gpu_fused_multi.sh <tag> <N>': decomposition tests + the bench at N GPUs, fused step on / off
set -u
TAG="${1:-fm}"; N="${2:-2}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
if [ "${SKIPTEST:-0}" != "1" ]; then timeout 900 python -m pytest tests/test_gpu_decomposition.py -m gpu -q -x > "$OUT/pytest.log" 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)" | tee -a "$OUT/summary.txt"; fi
for f in ${FUSED:-0 1}; do
  LS1B200_FUSED=$f timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_${N}_f$f.json" 2> "$OUT/bench_${N}_f$f.err"
  python - "$OUT/bench_${N}_f$f.json" "$N GPUs fused=$f" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-18s MMUPS %8.1f  ms/step %.3f  phases %s trace %s" % (sys.argv[2], d["value"], d["ms_per_step"], d["phases_ms_per_step"], d.get("trace")))
except Exception as e:
    print(sys.argv[2], "no bench line:", e)
PY
done
