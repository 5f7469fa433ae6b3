#!/bin/bash
# Standard measurement set of a round (run under gpurun, 1 GPU): tests, bench (both arms), ncu launch list, ncu full profile.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 600 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 2500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
tail -c 600 gpurun_out/bench_${TAG}_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 50 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"expand_kernel|walk_packs_parallel|msd_partition|msd_count|leaf_warp|leaf_gather|leaf_scan" -s 24 -c 8 -o gpurun_out/prof_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -12
