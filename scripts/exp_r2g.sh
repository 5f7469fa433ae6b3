#!/bin/bash
# round-2 experiment call 1: primitives + leaf knobs (probe_bin per-stage times)
set -u
mkdir -p gpurun_out
./scripts/ubench/ubench > gpurun_out/ubench_r2g.txt 2>&1
G7=build/libkmc_b200_g7.so
timeout 900 python scripts/sweep_env.py 117440512 31 "" "KMCB200_L2_BITS=10" "KMCB200_LIB=$G7" "KMCB200_LIB=$G7,KMCB200_LEAF_ROUND_PCT=150" "KMCB200_LIB=$G7,KMCB200_LEAF_ROUND_PCT=200" "KMCB200_LIB=$G7,KMCB200_LEAF_ROUND_PCT=350" 2>&1 | tee gpurun_out/sweep_r2g_117M.txt
timeout 600 python scripts/sweep_env.py 67108864 31 "" "KMCB200_LIB=$G7,KMCB200_LEAF_ROUND_PCT=200" 2>&1 | tee gpurun_out/sweep_r2g_64M.txt
timeout 300 python scripts/sweep_env.py 33554432 31 "" 2>&1 | tee gpurun_out/sweep_r2g_32M.txt
cat gpurun_out/ubench_r2g.txt
