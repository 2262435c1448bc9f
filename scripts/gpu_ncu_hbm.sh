#!/usr/bin/env bash
# gpurun -- 'bash scripts/gpu_ncu_hbm.sh <tag>': full ncu capture of the HBM-bound kernels (binning, integration) at the bench size
set -u
TAG="${1:-h}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gather|k_key|k_scatter|k_leapfrog_pre_v2|k_leapfrog_post_v2' -s 10 -c 5 -o "$OUT/hbm_full" -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > "$OUT/ncu_hbm.log" 2>&1
tail -2 "$OUT/ncu_hbm.log" | cut -c1-200
