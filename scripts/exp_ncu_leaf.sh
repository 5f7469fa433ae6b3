#!/bin/bash
# one ncu --set full capture of the leaf kernel (bin of 1.17e8 k-mers), with source counters; usage: exp_ncu_leaf.sh TAG [env assignments...]
set -u
mkdir -p gpurun_out
TAG=$1; shift
for kv in "$@"; do export "$kv"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"leaf_hash_kernel|leaf_warp_kernel" -s 2 -c 1 -o gpurun_out/prof_leaf_${TAG} -f python scripts/probe_bin.py 117440512 31 2 > gpurun_out/ncu_leaf_${TAG}.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_leaf_${TAG}.log; ls -la gpurun_out/prof_leaf_${TAG}.ncu-rep
