#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:edge_bwd_tc -s 7 -c 1 -o gpurun_out/prof_bwd -f $CMD > gpurun_out/ncu_bwd.log 2>&1
echo "bwd capture rc=$?"
