#!/bin/bash
# Standard measurement set of a round (run under gpurun, 1 GPU): tests, bench (both arms, the driver's flags), ncu launch list of the bench
# command, ncu --set full of every hot kernel on one bin of the workload's mean size.  usage: gpu_measure.sh TAG [notests]
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
if [ "${2:-}" != "notests" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_${TAG}.json | head -c 1500; echo; tail -3 gpurun_out/bench_${TAG}.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
echo "reference rc=$?"; tail -c 900 gpurun_out/bench_${TAG}_reference.json; echo
# launch list: 420 launches (= 20 bins) out of the timed region of the same command (3 warm-up steps + pool pass = ~43 K launches skipped)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 45000 -c 420 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/ncu_launches_${TAG}.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"expand_kernel|walk_packs_parallel|msd_partition|msd_count|leaf_warp" -s 6 -c 6 -o gpurun_out/prof_${TAG} python scripts/probe_bin.py 117440512 31 2 > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out | tail -8
