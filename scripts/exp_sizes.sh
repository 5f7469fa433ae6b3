#!/bin/bash
# per-stage times over the bin sizes of the workload's pool, for several settings.  usage: exp_sizes.sh TAG "cfg" "cfg" ...
set -u
mkdir -p gpurun_out
TAG=$1; shift
for n in 33554432 67108864 100663296 117440512 134217728 167772160 201326592 268435456; do
  echo "== $n" | tee -a gpurun_out/sizes_${TAG}.txt
  timeout 600 python scripts/sweep_env.py $n 31 "$@" 2>&1 | tee -a gpurun_out/sizes_${TAG}.txt
done
