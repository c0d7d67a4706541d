#!/bin/bash
# ncu evidence for one training step (run on the GPU box): launch list with device times, then full captures of the
# two tensor-core edge kernels.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -c 600 gpurun_out/plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:edge_bwd2 -s 7 -c 1 -o gpurun_out/prof_bwd -f $CMD > gpurun_out/ncu_bwd.log 2>&1
echo "bwd capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:edge_fwd2 -s 7 -c 1 -o gpurun_out/prof_fwd -f $CMD > gpurun_out/ncu_fwd.log 2>&1
echo "fwd capture rc=$?"
ls -la gpurun_out/*.ncu-rep
