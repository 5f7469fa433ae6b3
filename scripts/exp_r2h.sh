#!/bin/bash
# round-2 experiment call: leaf_hash_kernel - parity tests, then per-stage times against leaf_warp_kernel
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/tests_r2h.txt
timeout 600 python scripts/sweep_env.py 117440512 31 "" "KMCB200_LEAF_KERNEL=warp" "KMCB200_LEAF_FILL_PCT=50" "KMCB200_LEAF_FILL_PCT=75" "KMCB200_L2_BITS=9" 2>&1 | tee gpurun_out/sweep_r2h_117M.txt
timeout 600 python scripts/sweep_env.py 67108864 31 "" "KMCB200_LEAF_KERNEL=warp" 2>&1 | tee gpurun_out/sweep_r2h_64M.txt
timeout 600 python scripts/sweep_env.py 268435456 31 "" "KMCB200_LEAF_KERNEL=warp" "KMCB200_L2_BITS=10" 2>&1 | tee gpurun_out/sweep_r2h_256M.txt
