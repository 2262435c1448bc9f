#!/usr/bin/env bash
# gpurun --gpus N -- 'bash scripts/gpu_ov2.sh <tag> <N>': 2-GPU transport tests + the bench at N GPUs, exchange overlapped with the inner traversal on / off
set -u
TAG="${1:-ov2}"; N="${2:-2}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
if [ "${SKIPTEST:-0}" != "1" ]; then
timeout 900 python -m pytest tests/test_gpu_decomposition.py -m gpu -q -x -k "two_gpus" > "$OUT/pytest.log" 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)" | tee -a "$OUT/summary.txt"
tail -30 "$OUT/pytest.log" | cut -c1-400
fi
for f in ${OV2:-1 0}; do
  LS1B200_OVERLAP2=$f timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_${N}_o$f.json" 2> "$OUT/bench_${N}_o$f.err"
  python - "$OUT/bench_${N}_o$f.json" "$N GPUs overlap2=$f" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-18s MMUPS %8.1f  ms/step %.3f  phases %s xchg %.3f trace %s" % (sys.argv[2], d["value"], d["ms_per_step"], d["phases_ms_per_step"], d.get("exchange_overlapped_ms_per_step", -1), d.get("trace")))
except Exception as e:
    print(sys.argv[2], "no bench line:", e)
PY
  tail -3 "$OUT/bench_${N}_o$f.err"
done
