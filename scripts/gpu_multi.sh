#!/usr/bin/env bash
# gpurun --gpus N -- 'bash scripts/gpu_multi.sh <tag> <N>': decomposition tests incl. NCCL, weak-scaling bench lines at 1..N GPUs
set -u
TAG="${1:-m}"; NG="${2:-2}"
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt" 2>&1
timeout 900 python -m pytest tests/test_gpu_decomposition.py -m gpu -q > "$OUT/pytest_decomp.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"
tail -30 "$OUT/pytest_decomp.log"
for n in 1 2 4 8; do
  [ "$n" -gt "$NG" ] && break
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 30 --warmup 5 > "$OUT/bench_n$n.json" 2> "$OUT/bench_n$n.err"
  fi
  echo "bench n=$n rc=$?" | tee -a "$OUT/summary.txt"
  cut -c1-900 "$OUT/bench_n$n.json"; tail -5 "$OUT/bench_n$n.err"
done
