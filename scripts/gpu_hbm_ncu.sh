#!/usr/bin/env bash
# ncu --set full of the HBM-bound kernels of the fused step (one launch each) + the new tests
set -u
TAG="${1:-hbm}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_step_switches or kernel_variants" > "$OUT/pytest.log" 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
timeout 600 ncu --set full --clock-control none -k regex:"k_pre_key|k_gather|k_scatter|k_scan_add|k_chunks" -s 24 -c 6 -o "$OUT/hbm_full" -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > "$OUT/ncu.log" 2>&1
ls -la "$OUT"
