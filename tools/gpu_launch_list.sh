#!/bin/bash
# ncu launch list of one steady-state train step (per-kernel device time; cold-cache, serialised: compare SHARES)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-200
wc -l gpurun_out/launches.csv
