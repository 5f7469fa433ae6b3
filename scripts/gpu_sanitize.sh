#!/bin/bash
# compute-sanitizer memcheck + racecheck over the small parity tests of the round's new kernels (run under gpurun).  usage: gpu_sanitize.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
SEL="leaf_hash_cutoff or all_T or edge_bins or one_byte or bad_packs or (leaf_hash_round and distinct and 55)"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_memcheck_${TAG}.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_${TAG}.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_racecheck_${TAG}.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_${TAG}.log
