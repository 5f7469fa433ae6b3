#!/bin/bash
# quick check of a leaf-kernel change: the leaf / parity tests, then per-stage times.  usage: exp_quick.sh TAG ["ENV=.. sweep cfgs" ...]
set -u
mkdir -p gpurun_out
TAG=$1; shift
timeout 900 python -m pytest tests -m gpu -x -q -k "leaf or parity or golden or cutoff or distinct or prefix or heavy or dominant or block" 2>&1 | tail -4 | tee gpurun_out/tests_${TAG}.txt
timeout 600 python scripts/sweep_env.py 117440512 31 "$@" 2>&1 | tee gpurun_out/sweep_${TAG}_117M.txt
