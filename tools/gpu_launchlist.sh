#!/bin/bash
# ncu launch list of the bench command (per-launch device times; cold-cache and serialised: compare SHARES)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
