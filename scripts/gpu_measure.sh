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
# launch list: the first 840 launches of the same command = the untimed pass over the 8 pool bins (every bin size of the workload once, 21
# launches each) + the first ~32 bins of the first warm-up step; --kill: the remaining ~10^5 launches of the run are not replayed under ncu
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 840 --kill on --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/ncu_launches_${TAG}.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"expand_kernel|walk_packs_parallel|msd_partition|msd_count|leaf_hash" -s 6 -c 6 -o gpurun_out/prof_${TAG} python scripts/probe_bin.py 117440512 31 2 > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out | tail -8
