#!/usr/bin/env bash
# One gpurun call: smoke, parity tests, the bench line (both arms), ncu launch list and one full capture of the dominant kernel.
# Usage (from the repo root): gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh <tag>'
set -u
TAG="${1:-r02}"
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > "$OUT/cpu.txt" 2>&1
python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
tail -14 "$OUT/pytest_gpu.log"
timeout 900 python bench.py --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
cut -c1-1500 "$OUT/bench.json"; tail -3 "$OUT/bench.err"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "bench ref rc=$?" | tee -a "$OUT/summary.txt"; cat "$OUT/bench_ref.json"
timeout 600 python bench.py --precision double --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > "$OUT/bench_double.json" 2> "$OUT/bench_double.err"; echo "bench double rc=$?" | tee -a "$OUT/summary.txt"
# launch list of the same command (shares, not absolutes), reduced steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > "$OUT/bench_under_ncu.log" 2>&1
# full capture of the force kernel at the bench size (16 M molecules: dram bytes per launch = roofline.traffic)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_1clj_tiled -s 2 -c 1 -o "$OUT/force_full" -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > "$OUT/ncu_full.log" 2>&1
ls -la "$OUT"
