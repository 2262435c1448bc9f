#!/usr/bin/env bash
# gpurun --gpus N -- 'bash scripts/gpu_overlap2.sh <tag> <N>': multi-GPU parity (both transports) and the bench with the exchange overlapped / in line
set -u
TAG="${1:-ov}"; N="${2:-2}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_decomposition.py -m gpu -q -x > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"; tail -5 "$OUT/pytest.log"
LS1B200_OVERLAP=0 timeout 600 python -m pytest tests/test_gpu_decomposition.py -m gpu -q -x -k nccl_two > "$OUT/pytest_noov.log" 2>&1; echo "pytest (overlap off) rc=$?" | tee -a "$OUT/summary.txt"
for ov in 1 0; do
  LS1B200_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > "$OUT/bench_${N}_ov$ov.json" 2> "$OUT/bench_${N}_ov$ov.err"
  python - "$OUT/bench_${N}_ov$ov.json" "$N GPUs overlap=$ov" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-22s MMUPS %8.1f  ms/step %.3f  phases %s  exchange(overlapped) %.3f  clocks %s" % (sys.argv[2], d["value"], d["ms_per_step"], d["phases_ms_per_step"], d.get("exchange_overlapped_ms_per_step", -1), d["clocks"]))
except Exception as e:
    print(sys.argv[2], "no bench line:", e)
PY
  tail -3 "$OUT/bench_${N}_ov$ov.err"
done
