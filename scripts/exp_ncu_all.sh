#!/bin/bash
# ncu --set full of every hot kernel on one bin of 1.17e8 k-mers, with source counters; usage: exp_ncu_all.sh TAG [env assignments...]
set -u
mkdir -p gpurun_out
TAG=$1; shift
for kv in "$@"; do export "$kv"; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"expand_kernel|walk_packs_parallel|msd_partition|msd_count|leaf_hash|leaf_warp" -s 6 -c 6 -o gpurun_out/prof_all_${TAG} -f python scripts/probe_bin.py 117440512 31 2 > gpurun_out/ncu_all_${TAG}.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_all_${TAG}.log; ls -la gpurun_out/prof_all_${TAG}.ncu-rep
