#!/usr/bin/env bash
# gpurun --gpus N -- 'bash scripts/gpu_overlap_bench.sh <tag> <N>': the bench at N GPUs with the exchange overlapped / in line
set -u
TAG="${1:-ovb}"; N="${2:-2}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for ov in ${OVS:-1 0}; do
  LS1B200_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_${N}_ov$ov.json" 2> "$OUT/bench_${N}_ov$ov.err"
  python - "$OUT/bench_${N}_ov$ov.json" "$N GPUs overlap=$ov" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-22s MMUPS %8.1f  ms/step %.3f  phases %s  exchange(overlapped) %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], d["phases_ms_per_step"], d.get("exchange_overlapped_ms_per_step", -1)))
except Exception as e:
    print(sys.argv[2], "no bench line:", e)
PY
done
